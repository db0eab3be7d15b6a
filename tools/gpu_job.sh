# tools/gpu_job.sh -- one gpurun call of the round-1 measurement loop (edit per experiment)
set -x
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matvec") > gpurun_out/pytest_mv.log 2>&1; tail -3 gpurun_out/pytest_mv.log
python tools/sweep.py --nx 256 --reps 20 --configs 10,11,12,13,14,15,16,17,18,19,20,21 --flags 1 --line-tiles=0 > gpurun_out/sweep3.log 2>&1; cut -c1-160 gpurun_out/sweep3.log
ncu --set full --clock-control none --import-source on -k regex:'k_dot_maps|k_dot_walk|k_xr_update|k_chunk_prefix' -s 8 -c 8 -o gpurun_out/prof_dot_r01 python tools/run_cg.py --nx 256 --iters 6 > gpurun_out/ncu_dot.log 2>&1
tail -3 gpurun_out/ncu_dot.log
