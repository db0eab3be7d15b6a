set -x
nvidia-smi -L
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; cat gpurun_out/bench1.json
python tools/sweep.py --nx 256 --reps 20 > gpurun_out/sweep1.log 2>&1; tail -40 gpurun_out/sweep1.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python tools/run_cg.py --nx 256 --iters 10 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 2 -c 2 -o gpurun_out/prof_matvec_r01 python tools/run_cg.py --nx 256 --iters 5 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
