set -x
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dot or cg") > gpurun_out/pytest_dot.log 2>&1; tail -3 gpurun_out/pytest_dot.log
python tools/sweep.py --nx 256 --reps 20 --configs 15 --flags 1 --line-tiles=0 > gpurun_out/sweep4.log 2>&1; cut -c1-160 gpurun_out/sweep4.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cat gpurun_out/bench2.json
ncu --set full --clock-control none --import-source on -k regex:'k_dot_maps2' -s 4 -c 2 -o gpurun_out/prof_dot2_r01 python tools/run_cg.py --nx 256 --iters 6 > gpurun_out/ncu_dot2.log 2>&1
