#!/bin/bash
# Round-2 closing evidence run on ONE B200 (after the stride-2 / weight-gradient / bit-mask work): tests, smoke, bench, conv
# micro-benchmark, one ncu capture of the paired weight gradient, ncu launch list of an eager D + G pair -> gpurun_out/
mkdir -p gpurun_out
echo "==== full tests"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2g_tests.log 2>&1; tail -6 gpurun_out/r2g_tests.log
echo "==== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "==== BENCH (default flags)"
SAE_BENCH_CONV_TABLE=gpurun_out/r2g_conv_table.txt timeout 900 python bench.py 2>gpurun_out/r2g_bench.err | tail -1 > gpurun_out/r2g_bench.json; cut -c1-300 gpurun_out/r2g_bench.json
echo "==== BENCH --steps 20 --warmup 3 (the driver's flags)"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2g_bench20.json; cut -c1-200 gpurun_out/r2g_bench20.json
echo "==== conv micro-benchmark"
timeout 300 python scripts/conv_bench.py > gpurun_out/r2g_conv_bench.txt 2>&1; tail -56 gpurun_out/r2g_conv_bench.txt
echo "==== ncu full: paired weight gradient, stride 1"
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 600 $NCU -k regex:"wgrad_tc_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_wgrad_pair_256 python scripts/conv_bench.py --only "256->256 @128" --dirs wgrad --iters 1 > gpurun_out/r2_ncu_k.log 2>&1; tail -1 gpurun_out/r2_ncu_k.log
echo "==== ncu launch list (eager D + G half-steps)"
SAE_CUDA_GRAPHS=0 SAE_BENCH_MIN_WARM=2 SAE_BENCH_SKIP_R1_WARM=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 3300 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-context > gpurun_out/r2g_ncu_bench.log 2>&1
tail -1 gpurun_out/r2g_ncu_bench.log | cut -c1-150; wc -l gpurun_out/r2g_launches.csv
