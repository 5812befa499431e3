#!/bin/bash
# round 2, GPU call 1: validate the opt-in stride-2 kernels (conv_tc5m, conv_tc6), time them, baseline bench
mkdir -p gpurun_out
echo "==== experimental kernels: parity suite with the switches on"
SAE_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 1400 > gpurun_out/r2c1_exp.log 2>&1; tail -15 gpurun_out/r2c1_exp.log
echo "==== conv bench, stride-2 shapes"
timeout 200 python scripts/conv_bench.py --only s2 > gpurun_out/r2c1_cb_base.txt 2>&1; cat gpurun_out/r2c1_cb_base.txt
SAE_TC6=1 timeout 200 python scripts/conv_bench.py --only s2 --dirs fprop > gpurun_out/r2c1_cb_tc6.txt 2>&1; cat gpurun_out/r2c1_cb_tc6.txt
SAE_DGRAD_MERGED=1 timeout 200 python scripts/conv_bench.py --only s2 --dirs dgrad > gpurun_out/r2c1_cb_merged.txt 2>&1; cat gpurun_out/r2c1_cb_merged.txt
echo "==== bench base"
SAE_BENCH_CONV_TABLE=gpurun_out/r2c1_conv_table_base.txt timeout 500 python bench.py --no-cpu-baseline 2>gpurun_out/r2c1_bench_base.err | tail -1 > gpurun_out/r2c1_bench_base.json; cut -c1-250 gpurun_out/r2c1_bench_base.json
echo "==== bench tc6+merged"
SAE_TC6=1 SAE_DGRAD_MERGED=1 SAE_BENCH_CONV_TABLE=gpurun_out/r2c1_conv_table_exp.txt timeout 500 python bench.py --no-cpu-baseline 2>gpurun_out/r2c1_bench_exp.err | tail -1 > gpurun_out/r2c1_bench_exp.json; cut -c1-250 gpurun_out/r2c1_bench_exp.json
echo "==== bench batched discriminator passes"
SAE_BATCH_D=1 timeout 500 python bench.py --no-cpu-baseline 2>gpurun_out/r2c1_bench_batchd.err | tail -1 > gpurun_out/r2c1_bench_batchd.json; cut -c1-250 gpurun_out/r2c1_bench_batchd.json
