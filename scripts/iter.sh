#!/bin/bash
mkdir -p gpurun_out
echo "==== conv tests (SAE_TC5 default)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "conv" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -15 | tee gpurun_out/tests_conv.log
if grep -q "failed\|Timeout\|error" gpurun_out/tests_conv.log; then echo "CONV TESTS FAILED"; export SAE_TC5=0; fi
for m in 0 1 2; do
echo "==== conv bench SAE_TC5=$m"
SAE_TC5=$m timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad --only "1x1" 2>&1 | tail -12
SAE_TC5=$m timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | tail -4
SAE_TC5=$m timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad --only "Dpatch" 2>&1 | tail -5
SAE_TC5=$m timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad --only "E 32" 2>&1 | tail -2
done
echo "==== conv bench SAE_TC5=2 KB=36"
SAE_TC5_KB=36 timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad --only "s2" 2>&1 | tail -8
echo "==== graph tests"
timeout 600 python -m pytest tests/test_gpu_graphs.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -25 | tee gpurun_out/tests_graphs.log
echo "==== all parity tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -8 | tee gpurun_out/tests.log
echo "==== bench (tc5 default)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_graphs.log 2>&1; grep -B2 -A25 "capture of the" gpurun_out/bench_graphs.log | head -40; tail -1 gpurun_out/bench_graphs.log | tee gpurun_out/bench_graphs.json | cut -c1-300
echo "==== bench tc5 off"
SAE_TC5=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_tc5off.json | cut -c1-250
echo "==== bench tc5=1"
SAE_TC5=1 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_tc5_1.json | cut -c1-250
