#!/bin/bash
# round 2, GPU call 10: modulated-conv tests, ncu captures of the new kernels, launch list of one D + one G half-step
mkdir -p gpurun_out
echo "==== modulated conv tests"
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q --timeout 600 -k "modulated" > gpurun_out/r2c10_mod.log 2>&1; tail -15 gpurun_out/r2c10_mod.log
echo "==== is the per-sample path on in the step? (launch counters by kernel name via torch profiler, eager G half-step)"
timeout 600 python - <<'PY' 2>&1 | tail -40
import os, sys, torch
sys.path.insert(0, os.getcwd())
from torch.profiler import ProfilerActivity, profile
import swapping_autoencoder_pytorch_b200 as S
opt = S.default_options(num_gpus=1, batch_size=32, crop_size=256)
torch.manual_seed(0)
model = S.create_model(opt); trainer = S.create_optimizer(opt, model)
x = torch.randn(32, 3, 256, 256, device="cuda").clamp(-1, 1)
for _ in range(4): trainer.train_one_step({"real_A": x}, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer.train_one_step({"real_A": x}, 0); trainer.train_one_step({"real_A": x}, 0); torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:70]
        r = rows.setdefault(name, [0.0, 0]); r[0] += ev.device_time_total; r[1] += 1
tot = sum(r[0] for r in rows.values())
print("D + G half-step (eager): total device time %.2f ms over %d kernels" % (tot / 1e3, sum(r[1] for r in rows.values())))
for name, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:38]:
    print("%7.2f ms %5.1f%% %5d  %s" % (t / 1e3, 100 * t / tot, n, name))
PY
echo "==== ncu: conv_tc6<256> (stride-2 fprop), torgb, crop, adam, filter_modulate, modulated wgrad"
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:"conv_tc6_kernel" -s 1 -c 1 -o gpurun_out/r2_prof_s2_fprop_tc6_256 python scripts/conv_bench.py --only "D 128->256 @257 s2" --dirs fprop --iters 1 > gpurun_out/r2_ncu_g.log 2>&1; tail -1 gpurun_out/r2_ncu_g.log
timeout 400 $NCU -k regex:"torgb_fwd_kernel|torgb_bwd_kernel|crop_gather_kernel|crop_gather_bwd_kernel|adam_kernel|fir_tma_kernel" -c 8 -o gpurun_out/r2_prof_train_ops python scripts/train_ops_bench.py > gpurun_out/r2_ncu_h.log 2>&1; tail -3 gpurun_out/r2_ncu_h.log
timeout 200 python scripts/train_ops_bench.py > gpurun_out/r2_train_ops_bench.txt 2>&1; cat gpurun_out/r2_train_ops_bench.txt
