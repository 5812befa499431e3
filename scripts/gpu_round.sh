#!/bin/bash
# One GPU-box round trip (outputs under gpurun_out/).
mkdir -p gpurun_out
echo ==== full tests
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -12
echo ==== COPY AUDIT
B=8 timeout 300 python scripts/copy_audit.py 2>&1 | tail -30
echo ==== MEM BENCH
timeout 300 python scripts/mem_bench.py 2>&1 | tee gpurun_out/mem_bench.txt | grep -E "fir|bias_act fwd|modulate \[" | head -24
echo ==== CONV BENCH
timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad 2>&1 | tee gpurun_out/conv_bench.txt | head -14
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
