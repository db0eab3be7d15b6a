#!/bin/bash
# tools/gpu_r02_n2.sh N -- multi-GPU validation + benches on N GPUs (run under gpurun --gpus N)
N=${1:-2}
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
BIG=""
for f in tests/golden/big_128x128x128_np$N.npz tests/golden/big_256x256x256_np$N.npz; do [ -f $f ] && BIG="$BIG,$f"; done
[ $N = 2 ] && BIG="$BIG,tests/golden/big_256x256x512_np2.npz"
[ $N = 4 ] && BIG="$BIG,tests/golden/big_256x512x512_np4.npz"
BIG=${BIG#,}
IMB=$(ls tests/golden/mpi_*_np${N}_imb*.npz 2>/dev/null | tr '\n' ',' | sed 's/,$//')
$TR --master-port 29511 tests/mgpu_check.py --sizes 2x2x2,8x8x8,5x3x7,12x10x9,20x30x40,64x64x64${IMB:+,$IMB} --big "$BIG" --big-modes peer+overlap,nccl > $O/r02_mgpu_check_n$N.log 2>&1
echo "mgpu_check rc=$?" >> $O/r02_mgpu_check_n$N.log
python -m pytest tests/test_gpu_dropin.py -q -k multi_rank > $O/r02_dropin_np_n$N.log 2>&1
echo "dropin rc=$?" >> $O/r02_dropin_np_n$N.log
$TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > $O/r02_bench_n${N}_weak.json 2> $O/r02_bench_n${N}_weak.err
[ $N = 2 ] && $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --overlap 2 > $O/r02_bench_n${N}_weak_overlap2.json 2> $O/r02_bench_n${N}_weak_overlap2.err
$TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --strong --nx 400 > $O/r02_strong_400_n$N.json 2> $O/r02_strong_400_n$N.err
$TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --strong --nx 512 > $O/r02_strong_512_n$N.json 2> $O/r02_strong_512_n$N.err
tail -3 $O/r02_mgpu_check_n$N.log; tail -2 $O/r02_dropin_np_n$N.log
