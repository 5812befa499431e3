#!/bin/bash
mkdir -p gpurun_out
echo "==== graph tests"
timeout 600 python -m pytest tests/test_gpu_graphs.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" > gpurun_out/tests_graphs_full.log; tail -5 gpurun_out/tests_graphs_full.log; grep -n "Error\|error" gpurun_out/tests_graphs_full.log | head -20
echo "==== all parity tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -8 | tee gpurun_out/tests.log
echo "==== bench default"
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs.json | cut -c1-200
echo "==== bench per-operator blocks"
SAE_FUSED_BLOCKS=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_unfused.json | cut -c1-200
echo "==== bench no fused fir+act"
SAE_FUSED_FIR_ACT=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graphs_nofiract.json | cut -c1-200
echo "==== op profile (eager, kernel table only)"
timeout 300 python scripts/op_profile.py > gpurun_out/op_profile.txt 2>&1; grep "sae::\|at::native\|Self C" gpurun_out/op_profile.txt | cut -c1-50,130-215 | head -45
