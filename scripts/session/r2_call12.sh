#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 > gpurun_out/r2c12_trainops.log 2>&1; tail -6 gpurun_out/r2c12_trainops.log
timeout 200 python scripts/train_ops_bench.py 2>&1 | grep crop
bash scripts/r2_final.sh
