#!/bin/bash
mkdir -p gpurun_out
echo "==== conv tests (persistent kernel)"
timeout 300 python -m pytest tests -m gpu -q --timeout 100 -k "conv or shadow" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -12
echo ==== CONV BENCH persistent
timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad 2>&1 | tee gpurun_out/conv_bench.txt | head -30
echo ==== CONV BENCH non-persistent
SAE_TC_PERSISTENT=0 timeout 300 python scripts/conv_bench.py --dirs fprop --only "s1" 2>&1 | head -9
echo ==== full tests
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -6
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1.json | cut -c1-330
