#!/bin/bash
# round 2, GPU call 22: wide odd channel counts padded onto the tcgen05 kernels: parity at the ffhq1024 option set, step times of the BASELINE configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -m gpu -q --timeout 900 > gpurun_out/r2c22_full.log 2>&1; tail -5 gpurun_out/r2c22_full.log
timeout 900 python scripts/config_sweep.py 2>&1 | grep -E "ok|NaN|FAILED" 
