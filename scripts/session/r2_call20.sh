#!/bin/bash
# round 2, GPU call 20: activation bit masks (conv epilogue / FIR+act forward write them, the activation backward passes read them)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 > gpurun_out/r2c20_trainops.log 2>&1; tail -5 gpurun_out/r2c20_trainops.log
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graphs.py -m gpu -q -x --timeout 900 > gpurun_out/r2c20_parity.log 2>&1; tail -5 gpurun_out/r2c20_parity.log
for v in 1 0 1 0; do
SAE_ACT_MASK=$v timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c20_bench_m$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c20_bench_m$v.json')); print('act_mask=$v', d['value'], d['cadence']['ms'], d['roofline_hbm']['achieved'] if 'roofline_hbm' in d else None)"
done
