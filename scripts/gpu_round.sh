#!/bin/bash
mkdir -p gpurun_out
echo "==== conv tests (narrow wgrad, strips)"
if timeout 300 python -m pytest tests -m gpu -q --timeout 100 -x -k "conv_fprop_dgrad_wgrad" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -8 | tee gpurun_out/conv_tests.log | grep -q " passed"; then
  if grep -q failed gpurun_out/conv_tests.log; then echo "CONV TESTS FAILED -> narrow wgrad off"; export SAE_WGRAD_NARROW=0; fi
else echo "CONV TESTS DID NOT PASS -> narrow wgrad off"; export SAE_WGRAD_NARROW=0; fi
cat gpurun_out/conv_tests.log
echo "==== all tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | grep -v "^E   *+\|^E  *where" | tail -12
echo "==== mem bench (TMA FIR)"
timeout 300 python scripts/mem_bench.py 2>&1 | tee gpurun_out/mem_bench.txt | grep -E "fir|bias_act bwd|modulate bwd|add_scale" | head -40
echo "==== mem bench (strip FIR)"
SAE_FIR_TMA=0 timeout 300 python scripts/mem_bench.py 2>&1 | grep -E "fir" | head -12
echo "==== conv bench dgrad split / no split"
timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | tail -4
SAE_TC_SPLIT=0 timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | tail -3
echo "==== conv bench wgrad narrow / old"
timeout 300 python scripts/conv_bench.py --dirs wgrad --only "32->32" 2>&1 | tail -2
SAE_WGRAD_NARROW=0 timeout 300 python scripts/conv_bench.py --dirs wgrad --only "32->32" 2>&1 | tail -2
echo "==== BENCH"
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1.json | cut -c1-330
echo "==== ncu narrow"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_tc4_kernel|wgrad_narrow_kernel|wgrad_tc_kernel" -s 1 -c 2 -o gpurun_out/prof_narrow -f python scripts/conv_bench.py --only "Dpatch 32" --dirs fprop,wgrad --iters 1 > gpurun_out/ncu_narrow.log 2>&1; tail -2 gpurun_out/ncu_narrow.log
echo "==== ncu fir"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fir_tma_kernel|bias_act_bwd" -s 17 -c 2 -o gpurun_out/prof_fir -f python scripts/mem_bench.py > gpurun_out/ncu_fir.log 2>&1; tail -2 gpurun_out/ncu_fir.log
