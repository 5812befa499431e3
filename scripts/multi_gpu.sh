#!/bin/bash
# N-GPU bench exactly as the driver launches it (one process per GPU, NCCL over NVLink).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 8 --warmup 4 2>&1 | tail -3 | tee gpurun_out/bench_n$N.json | cut -c1-600
