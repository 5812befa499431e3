#!/bin/bash
# Round-end evidence run on ONE B200: tests, smoke, bench (both arms), micro-benchmarks, ncu captures -> gpurun_out/
mkdir -p gpurun_out
echo "==== full tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | grep -v "^E   *+\|^E  *where" | tail -6 | tee gpurun_out/tests_final.log
echo "==== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "==== BENCH"
SAE_BENCH_CONV_TABLE=gpurun_out/conv_table.txt timeout 900 python bench.py --warmup 4 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-300
echo "==== BENCH eager (SAE_CUDA_GRAPHS=0)"
SAE_CUDA_GRAPHS=0 timeout 900 python bench.py --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_final_eager.json | cut -c1-200
echo "==== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-300
echo "==== conv bench"
timeout 300 python scripts/conv_bench.py > gpurun_out/conv_bench.txt 2>&1; tail -48 gpurun_out/conv_bench.txt
echo "==== mem bench"
timeout 300 python scripts/mem_bench.py > gpurun_out/mem_bench.txt 2>&1; tail -5 gpurun_out/mem_bench.txt
echo "==== op profile"
timeout 300 python scripts/op_profile.py > gpurun_out/op_profile.txt 2>&1; grep "sae::\|Self C" gpurun_out/op_profile.txt | cut -c1-50,130-215 | head -12
echo "==== ncu full: ModConv fwd shapes + wgrad"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_tc5_kernel|wgrad_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_tc5_512 -f python scripts/conv_bench.py --only "512->512 @64" --dirs fprop,wgrad --iters 1 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_tc5_kernel" -s 1 -c 1 -o gpurun_out/prof_tc5_128 -f python scripts/conv_bench.py --only "128->128 @256" --dirs fprop --iters 1 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_tc5_kernel" -s 1 -c 1 -o gpurun_out/prof_tc5_1x1 -f python scripts/conv_bench.py --only "D skip 128->256 @128" --dirs fprop --iters 1 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
echo "==== ncu full: FIR"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fir_tma_kernel|bias_act_bwd_kernel" -s 1 -c 3 -o gpurun_out/prof_fir -f python scripts/mem_bench.py > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
echo "==== ncu launch list (eager D + G half-steps)"
SAE_CUDA_GRAPHS=0 SAE_BENCH_MIN_WARM=2 SAE_BENCH_SKIP_R1_WARM=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 5200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-150; wc -l gpurun_out/launches_final.csv
