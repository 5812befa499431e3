#!/bin/bash
mkdir -p gpurun_out
echo "==== fir tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "fir or resblock or shadow or layers or networks" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -6
echo "==== mem bench fir"
timeout 300 python scripts/mem_bench.py 2>&1 | grep "fir4 down2\|fir4 up2\|fused" | head -20
echo "==== bench N=2 (graphs on)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/bench_n2.log 2>&1; grep -i "capture of the\|Error" gpurun_out/bench_n2.log | head -10; tail -1 gpurun_out/bench_n2.log | tee gpurun_out/bench_n2.json | cut -c1-700
echo "==== bench N=1"
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs.json | cut -c1-200
