#!/bin/bash
# end-of-round evidence run on one B200:  gpurun -- bash tools/gpu_final.sh [tag]
TAG=${1:-r01d}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if d.get("impl") == "reference":
        print(sys.argv[1], "REFERENCE GF %.2f" % d["value"], d["cpu_baseline"])
    else:
        print(sys.argv[1], "GF %.1f ms/iter %.4f pct8 %.2f e2e %.1f mv_ms %.4f mvfrac %.3f" % (d["value"], d["ms_per_cg_iteration"], d["pct_of_8TBs_per_gpu"], d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"]), d["cg_times"], d["clocks"], d["setup_ms"], d.get("cpu_baseline"))
except Exception as e:
    print("no json", sys.argv[1], e)
PY
}
timeout 600 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; show gpurun_out/bench_n1_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; show gpurun_out/bench_ref_$TAG.json
timeout 300 python bench.py --steps 3 --warmup 3 --nx 512 --no-cpu-baseline > gpurun_out/bench_n1_512_$TAG.json 2> gpurun_out/bench_n1_512_$TAG.err; show gpurun_out/bench_n1_512_$TAG.json
timeout 300 python bench.py --steps 3 --warmup 3 --nx 400 --strong --no-cpu-baseline > gpurun_out/bench_n1_400_$TAG.json 2> gpurun_out/bench_n1_400_$TAG.err; show gpurun_out/bench_n1_400_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/run_cg.py --nx 256 --iters 12 > gpurun_out/ncu_launch_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt 2>&1; cat gpurun_out/launch_summary_$TAG.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 3 -c 1 -o gpurun_out/prof_matvec_$TAG python tools/run_cg.py --nx 256 --iters 6 > gpurun_out/ncu_mv_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_dot_ -s 6 -c 4 -o gpurun_out/prof_dot_$TAG python tools/run_cg.py --nx 256 --iters 6 > gpurun_out/ncu_dot_$TAG.log 2>&1
for r in prof_matvec_$TAG prof_dot_$TAG; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -20
