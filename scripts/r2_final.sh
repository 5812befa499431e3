#!/bin/bash
# Round-2 evidence run on ONE B200: tests, smoke, bench (both arms), micro-benchmarks, ncu captures -> gpurun_out/
mkdir -p gpurun_out
echo "==== full tests"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2f_tests.log 2>&1; tail -6 gpurun_out/r2f_tests.log
echo "==== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "==== BENCH (default flags)"
SAE_BENCH_CONV_TABLE=gpurun_out/r2f_conv_table.txt timeout 900 python bench.py 2>gpurun_out/r2f_bench.err | tail -1 > gpurun_out/r2f_bench.json; cut -c1-300 gpurun_out/r2f_bench.json
echo "==== BENCH --steps 20 --warmup 3 (the driver's flags)"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2f_bench20.json; cut -c1-200 gpurun_out/r2f_bench20.json
echo "==== BENCH eager (SAE_CUDA_GRAPHS=0)"
SAE_CUDA_GRAPHS=0 timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2f_bench_eager.json; cut -c1-200 gpurun_out/r2f_bench_eager.json
echo "==== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r2f_bench_ref.json; cut -c1-300 gpurun_out/r2f_bench_ref.json
echo "==== micro benches"
timeout 300 python scripts/conv_bench.py > gpurun_out/r2f_conv_bench.txt 2>&1; tail -50 gpurun_out/r2f_conv_bench.txt
timeout 300 python scripts/mem_bench.py > gpurun_out/r2f_mem_bench.txt 2>&1; tail -5 gpurun_out/r2f_mem_bench.txt
timeout 300 python scripts/train_ops_bench.py > gpurun_out/r2f_train_ops_bench.txt 2>&1; cat gpurun_out/r2f_train_ops_bench.txt
echo "==== R1 profile"
timeout 600 python scripts/r1_profile.py > gpurun_out/r2f_r1_profile.txt 2>&1; head -12 gpurun_out/r2f_r1_profile.txt
echo "==== ncu full: dominant kernel + modulated wgrad"
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 600 $NCU -k regex:"conv_tc5_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_tc5_128 python scripts/conv_bench.py --only "128->128 @256" --dirs fprop --iters 1 > gpurun_out/r2_ncu_i.log 2>&1; tail -1 gpurun_out/r2_ncu_i.log
timeout 600 $NCU -k regex:"wgrad_tc_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_wgrad_128 python scripts/conv_bench.py --only "128->128 @256" --dirs wgrad --iters 1 > gpurun_out/r2_ncu_j.log 2>&1; tail -1 gpurun_out/r2_ncu_j.log
echo "==== ncu launch list (eager D + G half-steps)"
SAE_CUDA_GRAPHS=0 SAE_BENCH_MIN_WARM=2 SAE_BENCH_SKIP_R1_WARM=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 3300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-context > gpurun_out/r2_ncu_bench.log 2>&1
tail -1 gpurun_out/r2_ncu_bench.log | cut -c1-150; wc -l gpurun_out/r2_launches.csv
