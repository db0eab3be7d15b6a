set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/mgpu2.log 2>&1; tail -8 gpurun_out/mgpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
ncu --set full --clock-control none --import-source on -k regex:'k_matvec' -s 3 -c 2 -o gpurun_out/prof_matvec2_r01 python tools/run_cg.py --nx 256 --iters 5 > gpurun_out/ncu_mv2.log 2>&1
