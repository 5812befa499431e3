#!/bin/bash
# round 2, GPU call 4: fused crop / Adam / closed-form R1 tests, whole suite, bench
mkdir -p gpurun_out
echo "==== train-ops tests"
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 2>&1 | tail -25 | tee gpurun_out/r2c4_trainops.log
echo "==== whole suite"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_gpu_train_ops.py 2>&1 | tail -25 | tee gpurun_out/r2c4_tests.log
echo "==== bench"
SAE_BENCH_CONV_TABLE=gpurun_out/r2c4_conv_table.txt timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>gpurun_out/r2c4_bench.err | tail -1 > gpurun_out/r2c4_bench.json; cut -c1-300 gpurun_out/r2c4_bench.json; tail -3 gpurun_out/r2c4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c4_bench.json'))
print(d['value'], d['e2e']['value'], d.get('cadence'))
PY
echo "==== bench batched D"
SAE_BATCH_D=1 timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c4_bench_batchd.json; cut -c1-200 gpurun_out/r2c4_bench_batchd.json
