#!/bin/bash
# round 2, GPU call 5: train-ops + suite with full logs, tc7 / tc6<256> validation, conv bench of the stride-2 family
mkdir -p gpurun_out
echo "==== train-ops tests"
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 > gpurun_out/r2c5_trainops.log 2>&1; tail -12 gpurun_out/r2c5_trainops.log
echo "==== whole suite"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_train_ops.py > gpurun_out/r2c5_tests.log 2>&1; tail -12 gpurun_out/r2c5_tests.log
echo "==== alternative kernel selections (tc7 = SAE_DGRAD_MERGED=2)"
SAE_TEST_EXPERIMENTAL=1 timeout 2400 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 1500 -k "MERGED=2" > gpurun_out/r2c5_exp.log 2>&1; tail -8 gpurun_out/r2c5_exp.log
echo "==== conv bench stride 2"
timeout 200 python scripts/conv_bench.py --only s2 --dirs fprop,dgrad > gpurun_out/r2c5_cb_default.txt 2>&1; cat gpurun_out/r2c5_cb_default.txt
SAE_DGRAD_MERGED=2 timeout 200 python scripts/conv_bench.py --only s2 --dirs dgrad > gpurun_out/r2c5_cb_tc7.txt 2>&1; cat gpurun_out/r2c5_cb_tc7.txt
echo "==== bench"
SAE_BENCH_CONV_TABLE=gpurun_out/r2c5_conv_table.txt timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>gpurun_out/r2c5_bench.err | tail -1 > gpurun_out/r2c5_bench.json; cut -c1-200 gpurun_out/r2c5_bench.json
SAE_DGRAD_MERGED=2 SAE_BATCH_D=1 timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c5_bench_tc7_batchd.json; cut -c1-200 gpurun_out/r2c5_bench_tc7_batchd.json
