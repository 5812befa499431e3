#!/bin/bash
# One GPU-box round trip: parity tests, smoke, conv microbench, ncu captures, bench (outputs under gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -15
echo ==== SMOKE
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo ==== CONV BENCH
timeout 600 python scripts/conv_bench.py 2>&1 | tee gpurun_out/conv_bench.txt | tail -45
echo ==== NCU full
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/prof_convtc_fprop -f python scripts/conv_bench.py --only "128->128 @256" --dirs fprop --iters 1 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/prof_convtc_512 -f python scripts/conv_bench.py --only "512->512 @64" --dirs fprop --iters 1 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad -f python scripts/conv_bench.py --only "128->128 @256" --dirs wgrad --iters 1 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
echo ==== BENCH
timeout 900 python bench.py --steps 8 --warmup 3 2>&1 | tail -3
