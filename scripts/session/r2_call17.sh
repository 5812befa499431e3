#!/bin/bash
# round 2, GPU call 17: batched remainder strips + deeper window ring of the grouped stride-2 dgrad: parity, timing;
# ncu captures of the narrow-layer kernels (conv_tc5<32>, <64>), the grouped dgrad and the paired weight gradient
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "conv or adjoint or networks or layers or shadow" > gpurun_out/r2c17_conv.log 2>&1; tail -4 gpurun_out/r2c17_conv.log
echo "== dgrad s2"
timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | grep -E "s2"
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:"conv_tc5_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_tc5_32 python scripts/conv_bench.py --only "Dpatch 32->32" --dirs fprop --iters 1 > gpurun_out/r2_ncu_a.log 2>&1; tail -1 gpurun_out/r2_ncu_a.log
timeout 300 $NCU -k regex:"conv_tc5_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_tc5_64 python scripts/conv_bench.py --only "Dpatch 64->64" --dirs fprop --iters 1 > gpurun_out/r2_ncu_b.log 2>&1; tail -1 gpurun_out/r2_ncu_b.log
timeout 300 $NCU -k regex:"conv_tc5m_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_dgrad_grouped python scripts/conv_bench.py --only "D 128->256 @257 s2" --dirs dgrad --iters 1 > gpurun_out/r2_ncu_c.log 2>&1; tail -1 gpurun_out/r2_ncu_c.log
timeout 300 $NCU -k regex:"wgrad_tc_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_wgrad_pair python scripts/conv_bench.py --only "D 128->256 @257 s2" --dirs wgrad --iters 1 > gpurun_out/r2_ncu_d.log 2>&1; tail -1 gpurun_out/r2_ncu_d.log
for v in 1 2; do
timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c17_bench_$v.json; python -c "
import json; d=json.load(open('gpurun_out/r2c17_bench_$v.json')); print('run $v', d['value'], d['cadence']['ms'], d['roofline']['achieved'])"
done
