#!/bin/bash
mkdir -p gpurun_out
echo "==== all tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | grep -v "^E   *+\|^E  *where" | tail -12 | tee gpurun_out/tests.log
echo "==== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "==== BENCH"
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 32 --warmup 4 2>&1 | tail -1 | tee gpurun_out/bench_r1.json | cut -c1-400
echo "==== mem bench"
timeout 300 python scripts/mem_bench.py 2>&1 | tee gpurun_out/mem_bench.txt | tail -40
echo "==== conv bench"
timeout 300 python scripts/conv_bench.py 2>&1 | tee gpurun_out/conv_bench.txt | tail -50
echo "==== NCU launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 5200 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "==== host floor: batch 2"
timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --per-gpu-batch 2 2>&1 | tail -1 | cut -c1-200
echo "==== copy audit"
B=8 timeout 300 python scripts/copy_audit.py 2>&1 | tail -30
echo "==== op profile"
timeout 300 python scripts/op_profile.py > gpurun_out/op_profile.txt 2>&1; grep -v "^---" gpurun_out/op_profile.txt | grep -i "aten::\|at::native\|Self C" | cut -c1-200 | head -50
