#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 > gpurun_out/r2c11_trainops.log 2>&1; tail -15 gpurun_out/r2c11_trainops.log
timeout 200 python scripts/train_ops_bench.py > gpurun_out/r2_train_ops_bench.txt 2>&1; cat gpurun_out/r2_train_ops_bench.txt
timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c11_bench.json; cut -c1-200 gpurun_out/r2c11_bench.json
