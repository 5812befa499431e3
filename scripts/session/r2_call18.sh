#!/bin/bash
# round 2, GPU call 18: two epilogue warpgroups (conv_tc5m<2>, conv_tc6<256>), 5-D tensor maps in the paired weight gradient
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or adjoint or networks or layers or shadow or train_steps" > gpurun_out/r2c18_conv.log 2>&1; tail -4 gpurun_out/r2c18_conv.log
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 -k "modulated" > gpurun_out/r2c18_trainops.log 2>&1; tail -3 gpurun_out/r2c18_trainops.log
echo "== s2 family"
timeout 300 python scripts/conv_bench.py --only "s2" 2>&1 | grep -E "s2"
echo "== wgrad, SAE_WGRAD_5D=0"
SAE_WGRAD_5D=0 timeout 300 python scripts/conv_bench.py --dirs wgrad --only "s" 2>&1 | grep -E "D/G|D 5|D 1|D 2"
echo "== wgrad, 5-D maps"
timeout 300 python scripts/conv_bench.py --dirs wgrad --only "s" 2>&1 | grep -E "D/G|D 5|D 1|D 2"
for v in 1 2; do
timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c18_bench_$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c18_bench_$v.json')); print('run $v', d['value'], d['cadence']['ms'], d['roofline']['achieved'], d['clocks'])"
done
