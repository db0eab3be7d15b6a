#!/bin/bash
# multi-GPU check + bench on N GPUs of one box:  gpurun --gpus N -- bash tools/gpu_mgpu.sh N [sizes]
N=${1:-2}
SIZES=${2:-8x8x8,20x30x40,64x64x64}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 $TR --master-port 29511 tests/mgpu_check.py --sizes $SIZES > gpurun_out/mgpu_check_n$N.log 2>&1
echo "mgpu_check rc=$?" | tee -a gpurun_out/mgpu_check_n$N.log
grep -E "MGPU_CHECK|MISMATCH|size " gpurun_out/mgpu_check_n$N.log | tail -40
for mode in "--comm peer" "--comm peer --no-overlap" "--comm nccl"; do
  tag=$(echo $mode | tr -d ' -')
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline $mode > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().splitlines()[-1])
    print("$mode", "GF", round(d["value"], 1), "ms/iter", round(d["ms_per_cg_iteration"], 4), "pct8", round(d["pct_of_8TBs_per_gpu"], 2), "e2e", round(d["e2e"]["value"], 1), d["config"].get("comm"))
except Exception as e:
    print("no json for $mode:", e)
PY
done
