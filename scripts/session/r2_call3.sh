#!/bin/bash
# round 2, GPU call 3: full GPU suite incl. the new BASELINE-config parity tests, new bench line (R1 cadence, context legs)
mkdir -p gpurun_out
echo "==== new parity tests"
timeout 1500 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_surface.py -m gpu -q --timeout 900 2>&1 | tail -25 | tee gpurun_out/r2c3_newtests.log
echo "==== bench"
SAE_BENCH_CONV_TABLE=gpurun_out/r2c3_conv_table.txt timeout 900 python bench.py 2>gpurun_out/r2c3_bench.err | tail -1 > gpurun_out/r2c3_bench.json; cut -c1-400 gpurun_out/r2c3_bench.json; tail -5 gpurun_out/r2c3_bench.err
echo "==== bench steps 20"
timeout 900 python bench.py --steps 20 --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c3_bench20.json; cut -c1-300 gpurun_out/r2c3_bench20.json
echo "==== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r2c3_bench_ref.json | cut -c1-400
echo "==== old suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_parity_full.py --deselect tests/test_gpu_surface.py 2>&1 | tail -6 | tee gpurun_out/r2c3_tests.log
