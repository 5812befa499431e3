#!/bin/bash
# round 2, GPU call 13: resident filter slice (narrow layers / 1x1 convs): parity + A/B timing on ONE box; per-phase step timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 > gpurun_out/r2c13_trainops.log 2>&1; tail -4 gpurun_out/r2c13_trainops.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or networks or layers or train_steps or shadow" > gpurun_out/r2c13_conv.log 2>&1; tail -4 gpurun_out/r2c13_conv.log
for v in 1 0; do
echo "== SAE_TC_RESIDENT_B=$v"
SAE_TC_RESIDENT_B=$v timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad 2>&1 | grep -E "1x1|Dpatch|E 32|FromRGB"
done
for v in 1 0 1 0; do
SAE_TC_RESIDENT_B=$v timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c13_bench_rb$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c13_bench_rb$v.json')); print('resident_b=$v', d['value'], d['cadence']['ms'], d.get('phases_rank0'))"
done
