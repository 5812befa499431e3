#!/bin/bash
mkdir -p gpurun_out
echo "==== full tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -6
echo "==== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "==== BENCH"
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 4 2>&1 | tail -1 | tee gpurun_out/bench_r1.json | cut -c1-400
echo "==== conv bench"
timeout 300 python scripts/conv_bench.py 2>&1 | tee gpurun_out/conv_bench.txt | tail -40
echo "==== op profile"
timeout 300 python scripts/op_profile.py > gpurun_out/op_profile.txt 2>&1; grep -v "^---" gpurun_out/op_profile.txt | cut -c1-230 | head -64
echo "==== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-300
echo "==== NCU launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 5200 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
