#!/bin/bash
# round 2, GPU call 9: per-sample-filter modulated conv, pruned conv file, whole suite, bench
mkdir -p gpurun_out
echo "==== train-ops tests (incl. modulated conv)"
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 > gpurun_out/r2c9_trainops.log 2>&1; tail -15 gpurun_out/r2c9_trainops.log
echo "==== whole suite"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_train_ops.py > gpurun_out/r2c9_tests.log 2>&1; tail -12 gpurun_out/r2c9_tests.log
echo "==== bench"
SAE_BENCH_CONV_TABLE=gpurun_out/r2c9_conv_table.txt timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>gpurun_out/r2c9_bench.err | tail -1 > gpurun_out/r2c9_bench.json; cut -c1-200 gpurun_out/r2c9_bench.json; tail -2 gpurun_out/r2c9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench.json'))
print(d['value'], d['e2e']['value'], d.get('cadence'))
PY
