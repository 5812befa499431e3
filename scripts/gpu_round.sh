#!/bin/bash
mkdir -p gpurun_out
echo ==== full tests
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -8
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1.json | cut -c1-330
echo ==== NCU launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 5000 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-120
