#!/bin/bash
# One GPU-box round trip (outputs under gpurun_out/).
mkdir -p gpurun_out
echo "==== conv tests with the CTA-pair kernel"
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -k "conv or shadow" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -15
echo ==== CONV BENCH pair
timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad 2>&1 | tee gpurun_out/conv_bench_pair.txt | tail -30
echo ==== CONV BENCH no pair
SAE_TC_PAIR=0 timeout 300 python scripts/conv_bench.py --dirs fprop --only "D/G" 2>&1 | tail -8
echo ==== full tests
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -8
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
