#!/bin/bash
# round 2, GPU call 15: stride-2 data gradient with two parity classes per work item (conv_tc5m<2>): parity, then A/B timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or adjoint or networks or layers or shadow" > gpurun_out/r2c15_conv.log 2>&1; tail -4 gpurun_out/r2c15_conv.log
for v in 2 1; do
echo "== SAE_DGRAD_MERGED=$v"
SAE_DGRAD_MERGED=$v timeout 300 python scripts/conv_bench.py --dirs dgrad --only s2 2>&1 | grep -E "s2"
SAE_DGRAD_MERGED=$v timeout 300 python scripts/conv_bench.py --dirs dgrad --only "@257 s2" --batch 80 2>&1 | grep -E "s2"
done
for v in 2 1 2 1; do
SAE_DGRAD_MERGED=$v timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c15_bench_m$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c15_bench_m$v.json')); print('merged=$v', d['value'], d['cadence']['ms'])"
done
