#!/bin/bash
# round 2, GPU call 16: weight gradient as CTA pairs (wgrad_tc_kernel<2>): parity, A/B timing; cost of the dgrad remainder strips
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or adjoint or networks or layers or shadow or train_steps" > gpurun_out/r2c16_conv.log 2>&1; tail -4 gpurun_out/r2c16_conv.log
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 -k "modulated" > gpurun_out/r2c16_trainops.log 2>&1; tail -4 gpurun_out/r2c16_trainops.log
for v in 1 0; do
echo "== SAE_WGRAD_PAIR=$v"
SAE_WGRAD_PAIR=$v timeout 300 python scripts/conv_bench.py --dirs wgrad 2>&1 | grep -E "s1|s2|1x1"
done
echo "== dgrad strips"
timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | grep -E "@25[67] s2"
for v in 1 0 1 0; do
SAE_WGRAD_PAIR=$v timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c16_bench_p$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c16_bench_p$v.json')); print('pair=$v', d['value'], d['cadence']['ms'])"
done
