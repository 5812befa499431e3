#!/bin/bash
mkdir -p gpurun_out
echo ==== CONV BENCH
timeout 300 python scripts/conv_bench.py --dirs fprop,dgrad --only "s1" 2>&1 | head -12
echo ==== full tests
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -6
echo ==== BENCH
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --steps 8 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1.json | cut -c1-330
echo ==== NCU full tc3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 1 -c 1 -o gpurun_out/prof_tc3_512 -f python scripts/conv_bench.py --only "512->512 @64" --dirs fprop --iters 1 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_kernel -s 1 -c 1 -o gpurun_out/prof_tc3_128 -f python scripts/conv_bench.py --only "128->128 @256" --dirs fprop --iters 1 > gpurun_out/ncu2.log 2>&1
ls gpurun_out | head -30
