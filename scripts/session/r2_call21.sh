#!/bin/bash
# round 2, GPU call 21: 5-D tensor maps for the narrow (32 / 64 input channel) weight gradients
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or adjoint or networks or layers or shadow or train_steps" > gpurun_out/r2c21_conv.log 2>&1; tail -4 gpurun_out/r2c21_conv.log
for v in 0 1; do
echo "== wgrad, SAE_WGRAD_5D=$v"
SAE_WGRAD_5D=$v timeout 300 python scripts/conv_bench.py --dirs wgrad 2>&1 | grep -E "Dpatch|E 32|FromRGB|D 128->256 @257"
done
for v in 1 2; do
timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c21_bench_$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c21_bench_$v.json')); print('run $v', d['value'], d['cadence']['ms'], d['roofline']['achieved'])"
done
