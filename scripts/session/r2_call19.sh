#!/bin/bash
# round 2, GPU call 19: 5-D maps in the single-CTA weight gradient; timing split of the grouped stride-2 data gradient
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or adjoint or networks or layers or shadow or train_steps" > gpurun_out/r2c19_conv.log 2>&1; tail -4 gpurun_out/r2c19_conv.log
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 -k "modulated" > gpurun_out/r2c19_trainops.log 2>&1; tail -3 gpurun_out/r2c19_trainops.log
echo "== wgrad"
timeout 300 python scripts/conv_bench.py --dirs wgrad 2>&1 | grep -E "wgrad"
for d in 0 1 2 4 6; do
echo "== dgrad s2, SAE_TC_DEBUG=$d"
SAE_TC_DEBUG=$d timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | grep -E "D 128|D 256|D 512"
done
for v in 1 2; do
timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c19_bench_$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c19_bench_$v.json')); print('run $v', d['value'], d['cadence']['ms'], d['roofline']['achieved'], d['clocks'])"
done
