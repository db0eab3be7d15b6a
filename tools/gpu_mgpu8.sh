#!/bin/bash
# N-GPU parity + weak/strong bench:  gpurun --gpus N -- bash tools/gpu_mgpu8.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "GF %.1f ms/iter %.4f pct8/gpu %.2f e2e %.1f mv_ms %.4f" % (d["value"], d["ms_per_cg_iteration"], d["pct_of_8TBs_per_gpu"], d["e2e"]["value"], d["roofline"]["ms_per_launch"]), d["config"].get("comm"), d["clocks"])
except Exception as e:
    print("no json", sys.argv[1], e)
PY
}
timeout 300 $TR --master-port 29511 tests/mgpu_check.py --sizes 8x8x8,64x64x64 > gpurun_out/mgpu_check_n$N.log 2>&1
echo "mgpu_check rc=$?" | tee -a gpurun_out/mgpu_check_n$N.log
grep -E "MGPU_CHECK|MISMATCH|size |Error|error" gpurun_out/mgpu_check_n$N.log | tail -20
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --comm peer > gpurun_out/bench_n${N}_peer_overlap.json 2> gpurun_out/bench_n${N}_peer_overlap.err; echo "rc=$?"; show gpurun_out/bench_n${N}_peer_overlap.json
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --comm peer --no-overlap > gpurun_out/bench_n${N}_peer.json 2> gpurun_out/bench_n${N}_peer.err; echo "rc=$?"; show gpurun_out/bench_n${N}_peer.json
timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --nx 400 --strong > gpurun_out/bench_n${N}_strong400.json 2> gpurun_out/bench_n${N}_strong400.err; echo "rc=$?"; show gpurun_out/bench_n${N}_strong400.json
if [ "${2:-}" = "nccl" ]; then
timeout 300 $TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --comm nccl > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?"; show gpurun_out/bench_n${N}_nccl.json
fi
tail -3 gpurun_out/bench_n${N}_peer_overlap.err
