#!/bin/bash
# round 2, GPU call 6: shadow + crop re-check, tc7 / tc5x validation and timing, R1 profile
mkdir -p gpurun_out
echo "==== re-check"
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "crop or shadow" > gpurun_out/r2c6_recheck.log 2>&1; tail -5 gpurun_out/r2c6_recheck.log
echo "==== tc5x (one CTA per SM, pooled rings): conv parity with SAE_TC5X=1 and 2"
for v in 1 2; do
SAE_TC5X=$v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "conv or networks or layers or train_steps or shadow" > gpurun_out/r2c6_tc5x$v.log 2>&1; tail -3 gpurun_out/r2c6_tc5x$v.log
done
echo "==== conv bench stride-1 wide shapes: default, TC5X=1, TC5X=2"
timeout 200 python scripts/conv_bench.py --only "s1" --dirs fprop,dgrad > gpurun_out/r2c6_cb_s1_default.txt 2>&1; cat gpurun_out/r2c6_cb_s1_default.txt
SAE_TC5X=1 timeout 200 python scripts/conv_bench.py --only "s1" --dirs fprop,dgrad > gpurun_out/r2c6_cb_s1_x1.txt 2>&1; cat gpurun_out/r2c6_cb_s1_x1.txt
SAE_TC5X=2 timeout 200 python scripts/conv_bench.py --only "s1" --dirs fprop,dgrad > gpurun_out/r2c6_cb_s1_x2.txt 2>&1; cat gpurun_out/r2c6_cb_s1_x2.txt
echo "==== tc7 parity"
SAE_DGRAD_MERGED=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "conv or layers or networks" > gpurun_out/r2c6_tc7.log 2>&1; tail -3 gpurun_out/r2c6_tc7.log
echo "==== bench TC5X=2"
SAE_TC5X=2 timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c6_bench_x2.json; cut -c1-200 gpurun_out/r2c6_bench_x2.json
SAE_TC5X=1 timeout 900 python bench.py --no-cpu-baseline --no-gpu-context 2>/dev/null | tail -1 > gpurun_out/r2c6_bench_x1.json; cut -c1-200 gpurun_out/r2c6_bench_x1.json
echo "==== R1 profile"
timeout 600 python scripts/r1_profile.py > gpurun_out/r2c6_r1_profile.txt 2>&1; head -45 gpurun_out/r2c6_r1_profile.txt
echo "==== ncu tc7"
SAE_DGRAD_MERGED=2 timeout 300 ncu --set full --import-source on --clock-control none -f -k regex:"conv_tc7_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_dgrad_tc7 python scripts/conv_bench.py --only "D 128->256 @257 s2" --dirs dgrad --iters 1 > gpurun_out/r2_ncu_f.log 2>&1; tail -1 gpurun_out/r2_ncu_f.log
