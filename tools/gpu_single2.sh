#!/bin/bash
TAG=${1:-r01c}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "GF %.1f ms/iter %.4f pct8 %.2f e2e %.1f mv_ms %.4f mvfrac %.3f" % (d["value"], d["ms_per_cg_iteration"], d["pct_of_8TBs_per_gpu"], d["e2e"]["value"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"]), d["cg_times"], d["clocks"])
except Exception as e:
    print("no json", sys.argv[1], e)
PY
}
for pdl in 0 1; do
  for ot in 2 0; do
    timeout 600 python bench.py --steps 5 --warmup 3 --pdl $pdl --op-timing $ot --no-cpu-baseline > gpurun_out/bench_n1_pdl${pdl}_ot${ot}_$TAG.json 2> gpurun_out/bench_n1_pdl${pdl}_ot${ot}_$TAG.err; show gpurun_out/bench_n1_pdl${pdl}_ot${ot}_$TAG.json
  done
done
timeout 600 python bench.py --steps 2 --warmup 3 --nx 400 --strong --no-cpu-baseline > gpurun_out/bench_n1_400_$TAG.json 2> gpurun_out/bench_n1_400_$TAG.err; show gpurun_out/bench_n1_400_$TAG.json
timeout 600 python bench.py --steps 2 --warmup 3 --nx 512 --no-cpu-baseline > gpurun_out/bench_n1_512_$TAG.json 2> gpurun_out/bench_n1_512_$TAG.err; show gpurun_out/bench_n1_512_$TAG.json; tail -3 gpurun_out/bench_n1_512_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/run_cg.py --nx 256 --iters 12 > gpurun_out/ncu_launch_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt 2>&1; cat gpurun_out/launch_summary_$TAG.txt
