#!/bin/bash
# N-GPU bench exactly as the driver launches it (one process per GPU, NCCL over NVLink).  usage: multi_gpu.sh N [tag] [extra env...]
N=${1:-2}
TAG=${2:-n$N}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 32 --warmup 4 > gpurun_out/bench_$TAG.log 2>&1
tail -1 gpurun_out/bench_$TAG.log > gpurun_out/bench_$TAG.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("$TAG", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["cuda_graphs"], d.get("strong"), d.get("cadence"))
except Exception as e:
    print("no json:", e)
    print(open("gpurun_out/bench_$TAG.log").read()[-3000:])
PY
