#!/usr/bin/env python
"""BASELINE.json configs[2..4] as run checks: one D (+R1) and one G half-step per configuration at a small batch."""
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swapping_autoencoder_pytorch_b200 as S  # noqa: E402

CONFIGS = {
    "ffhq512 (configs[2] shape)": dict(crop_size=512, batch_size=2),
    "ffhq1024 launcher options (configs[3])": dict(crop_size=1024, batch_size=2, netG_scale_capacity=0.8, netE_num_downsampling_sp=5,
                                                   netE_scale_capacity=0.4, global_code_ch=1536, patch_size=256),
    "1024 default nets": dict(crop_size=1024, batch_size=2),
    "afhq patch 32 (configs[4])": dict(crop_size=256, batch_size=4, patch_size=32),
    "afhq patch 64 (configs[4])": dict(crop_size=256, batch_size=4, patch_size=64),
    "church (no aggregation)": dict(crop_size=256, batch_size=4, patch_use_aggregation=False),
}
only = sys.argv[1:] or list(CONFIGS)
for name in only:
    over = CONFIGS[name]
    try:
        opt = S.default_options(num_gpus=1, R1_once_every=1, **over)
        torch.manual_seed(0)
        model = S.create_model(opt)
        trainer = S.create_optimizer(opt, model)
        x = torch.randn(opt.batch_size, 3, opt.crop_size, opt.crop_size, device="cuda").clamp(-1, 1)
        trainer.train_one_step({"real_A": x}, 0)
        trainer.train_one_step({"real_A": x}, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d = trainer.train_one_step({"real_A": x}, 0)
        g = trainer.train_one_step({"real_A": x}, 0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ok = all(float(v) == float(v) for v in list(d.values()) + list(g.values()))
        print("%-42s %s  D+R1+G %.0f ms  D_total %.3f G_L1 %.3f  peak mem %.1f GB" % (
            name, "ok " if ok else "NaN", dt * 1e3, float(d["D_total"]), float(g["G_L1"]),
            torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
        del model, trainer, x
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    except Exception as e:      # noqa: BLE001
        print("%-42s FAILED: %s: %s" % (name, type(e).__name__, str(e).splitlines()[0][:300]), flush=True)
        traceback.print_exc(limit=6)
