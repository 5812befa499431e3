#!/usr/bin/env python
"""Which host-level ops launch the torch (non-sae) kernels: torch.profiler over one D(+R1) and one G half-step."""
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
for _ in range(4):
    trainer.train_one_step({"real_A": x}, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    trainer.train_one_step({"real_A": x}, 0)
    trainer.train_one_step({"real_A": x}, 0)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=70,
                                                         max_name_column_width=48, max_shapes_column_width=70))
