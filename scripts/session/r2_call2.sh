#!/bin/bash
# round 2, GPU call 2: ncu --set full of the stride-2 conv family (fprop tc2/tc6, merged dgrad tc5m, wgrad)
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
S='D 128->256 @257 s2'
timeout 300 $NCU -k regex:"conv_tc2_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_fprop_tc2 python scripts/conv_bench.py --only "$S" --dirs fprop --iters 1 > gpurun_out/r2_ncu_a.log 2>&1; tail -1 gpurun_out/r2_ncu_a.log
SAE_TC6=1 timeout 300 $NCU -k regex:"conv_tc6_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_fprop_tc6 python scripts/conv_bench.py --only "$S" --dirs fprop --iters 1 > gpurun_out/r2_ncu_b.log 2>&1; tail -1 gpurun_out/r2_ncu_b.log
SAE_DGRAD_MERGED=1 timeout 300 $NCU -k regex:"conv_tc5m_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_dgrad_tc5m python scripts/conv_bench.py --only "$S" --dirs dgrad --iters 1 > gpurun_out/r2_ncu_c.log 2>&1; tail -1 gpurun_out/r2_ncu_c.log
timeout 300 $NCU -k regex:"wgrad_tc_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_wgrad python scripts/conv_bench.py --only "$S" --dirs wgrad --iters 1 > gpurun_out/r2_ncu_d.log 2>&1; tail -1 gpurun_out/r2_ncu_d.log
timeout 300 $NCU -k regex:"conv_tc5_kernel" -s 4 -c 4 -o gpurun_out/r2_prof_s2_dgrad_classes python scripts/conv_bench.py --only "$S" --dirs dgrad --iters 1 > gpurun_out/r2_ncu_e.log 2>&1; tail -1 gpurun_out/r2_ncu_e.log
ls -la gpurun_out/*.ncu-rep
