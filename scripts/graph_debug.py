#!/usr/bin/env python
"""Debug aid: capture the half-step graphs of the reduced model under a few option sets and print the full failure."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swapping_autoencoder_pytorch_b200 as S  # noqa: E402
from oracle.fixtures import TINY  # noqa: E402

warnings.simplefilter("always")
VARIANTS = {
    "full": {},
    "no patch D": dict(lambda_PatchGAN=0.0, lambda_patch_R1=0.0),
    "patch D, no patch R1": dict(lambda_patch_R1=0.0),
    "no R1 at all": dict(lambda_R1=0.0, lambda_patch_R1=0.0),
    "full, batch 4": dict(batch_size=4),
    "full, 128px": dict(crop_size=128, netE_num_downsampling_sp=4),
}
for name in (sys.argv[1:] or list(VARIANTS)):
    opt = S.default_options(**dict(TINY, num_gpus=1, cuda_graphs=True, R1_once_every=2, **VARIANTS[name]))
    torch.manual_seed(0)
    tr = S.create_optimizer(opt, S.create_model(opt))
    real = torch.randn(opt.batch_size, 3, opt.crop_size, opt.crop_size, device="cuda").clamp(-1, 1)
    g = tr.graphs
    t = tr
    for rnd_ in range(g.warmup + 1):
        for kind, body in (("D", t._discriminator_body), ("R1", t._r1_body), ("G", t._generator_body)):
            if kind == "R1" and opt.lambda_R1 == 0.0 and opt.lambda_patch_R1 == 0.0:
                continue
            g.run(kind, body, real)
            if g.disabled:
                break
        if g.disabled:
            break
    torch.cuda.synchronize()
    print("=== %-24s captured %s  disabled: %s" % (name, sorted(k[0] for k in g.captured), g.disabled), flush=True)
    if g.disabled:
        print(g.last_traceback[-3500:], flush=True)
