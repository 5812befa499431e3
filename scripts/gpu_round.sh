#!/bin/bash
# One GPU-box round trip (outputs under gpurun_out/).
mkdir -p gpurun_out
echo "==== wgrad shared-window mode: conv tests"
SAE_WGRAD_WINDOW=1 timeout 300 python -m pytest tests -m gpu -q --timeout 120 -k "conv" 2>&1 | grep -v "^E   *+\|^E  *where" | tail -12
echo ==== CONV BENCH wgrad window
SAE_WGRAD_WINDOW=1 timeout 300 python scripts/conv_bench.py --dirs wgrad --only "s1" 2>&1 | tail -8
echo ==== CONV BENCH wgrad plain
timeout 300 python scripts/conv_bench.py --dirs wgrad 2>&1 | tail -13
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo ==== NCU launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4000 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
