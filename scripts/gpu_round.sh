#!/bin/bash
# One GPU-box round trip: parity tests, smoke, conv microbench, bench with per-shape conv table (outputs under gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -15
echo ==== SMOKE
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo ==== CONV BENCH
timeout 600 python scripts/conv_bench.py 2>&1 | tee gpurun_out/conv_bench.txt | tail -45
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
head -50 gpurun_out/conv_table.txt
