#!/bin/bash
mkdir -p gpurun_out
echo "==== graph tests"
timeout 600 python -m pytest tests/test_gpu_graphs.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -25 | tee gpurun_out/tests_graphs.log
echo "==== parity tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -15 | tee gpurun_out/tests.log
for kb in 0 4 8 16 36; do
echo "==== conv bench SAE_TC_PERSIST_KB=$kb"
SAE_TC_PERSIST_KB=$kb timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad --only "1x1" 2>&1 | tail -12
SAE_TC_PERSIST_KB=$kb timeout 300 python scripts/conv_bench.py --dirs dgrad --only "s2" 2>&1 | tail -5
done
echo "==== bench graphs on (fused fir+act on)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_graphs.log 2>&1; grep -B2 -A25 "capture of the" gpurun_out/bench_graphs.log | head -60; tail -1 gpurun_out/bench_graphs.log | tee gpurun_out/bench_graphs.json | cut -c1-300
echo "==== bench graphs on, fused fir+act off"
SAE_FUSED_FIR_ACT=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_nofiract.json | cut -c1-250
echo "==== bench graphs on, persist kb 8 / 16"
SAE_TC_PERSIST_KB=8 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_kb8.json | cut -c1-250
SAE_TC_PERSIST_KB=16 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_kb16.json | cut -c1-250
