#!/bin/bash
# round 2, GPU call 8 (2 GPUs): NCCL all-reduce + Adam captured into the half-step graphs vs eager tail
bash scripts/multi_gpu.sh 2 n2_graphnccl
SAE_GRAPH_NCCL=0 bash scripts/multi_gpu.sh 2 n2_eagertail
grep -i "nccl\|error\|warn" gpurun_out/bench_n2_graphnccl.log | head -20
