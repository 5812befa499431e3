#!/usr/bin/env python
"""Where the lazy-R1 evaluation spends its device time: torch.profiler over the eager R1 body (batch 32, 256x256, default nets),
kernels grouped by name."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swapping_autoencoder_pytorch_b200 as S  # noqa: E402

B = int(os.environ.get("B", "32"))
opt = S.default_options(num_gpus=1, batch_size=B, crop_size=256)
torch.manual_seed(0)
model = S.create_model(opt)
trainer = S.create_optimizer(opt, model)
x = torch.randn(B, 3, 256, 256, device="cuda").clamp(-1, 1)
for _ in range(3):
    trainer._r1_body(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
trainer._r1_body(x)
e1.record()
torch.cuda.synchronize()
print("eager R1 body: %.2f ms" % e0.elapsed_time(e1))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    trainer._r1_body(x)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("<")[0].split("(")[0][:60]
        r = rows.setdefault(name, [0.0, 0])
        r[0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        r[1] += 1
tot = sum(r[0] for r in rows.values())
print("total device time %.2f ms over %d kernels" % (tot / 1e3, sum(r[1] for r in rows.values())))
for name, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%7.2f ms %5.1f%% %5d  %s" % (t / 1e3, 100 * t / tot, n, name))
