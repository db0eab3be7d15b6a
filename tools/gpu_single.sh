#!/bin/bash
# single-GPU regression + bench + profiles:  gpurun -- bash tools/gpu_single.sh [tag]
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi -L
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_n1_$TAG.json
timeout 900 python bench.py --steps 3 --warmup 3 --nx 400 --strong --no-cpu-baseline > gpurun_out/bench_n1_400_$TAG.json 2> gpurun_out/bench_n1_400_$TAG.err; echo "bench400 rc=$?"; cat gpurun_out/bench_n1_400_$TAG.json
timeout 900 python bench.py --steps 2 --warmup 3 --nx 512 --no-cpu-baseline > gpurun_out/bench_n1_512_$TAG.json 2> gpurun_out/bench_n1_512_$TAG.err; echo "bench512 rc=$?"; cat gpurun_out/bench_n1_512_$TAG.json; tail -3 gpurun_out/bench_n1_512_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python tools/run_cg.py --nx 256 --iters 12 > gpurun_out/ncu_launch_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt 2>&1; cat gpurun_out/launch_summary_$TAG.txt
