#!/usr/bin/env python
"""Probe for DESIGN.md open item 1 (one unexplained miss of the eager / CUDA-graph step comparison).

  python scripts/determinism_probe.py [repeats]

(1) forward determinism: E, G, D, Dpatch evaluated twice on identical inputs must agree bit for bit (no atomics on the
    forward path); (2) eager-vs-eager: two eager trainers, the second re-seeded with the first one's state before every
    half-step — if THIS drifts beyond kernel noise, the graph replay is not the culprit; (3) eager-vs-graph, the test's
    own comparison, with every step's numbers printed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import swapping_autoencoder_pytorch_b200 as S  # noqa: E402
from oracle.fixtures import TINY  # noqa: E402
from swapping_autoencoder_pytorch_b200.stylegan2_layers import NoiseInjection  # noqa: E402


def zero_noise(self, image, noise=None):
    if self.image_size is None:
        self.image_size = image.shape
    b, _, h, w = image.shape
    return image.new_empty(b, 1, h, w).zero_()


NoiseInjection.resolve_noise = zero_noise
DET = dict(lambda_PatchGAN=0.0, lambda_patch_R1=0.0, R1_once_every=2)


def trainer(graphs, **over):
    opt = S.default_options(**dict(TINY, num_gpus=1, cuda_graphs=graphs, **DET, **over))
    torch.manual_seed(0)
    return S.create_optimizer(opt, S.create_model(opt))


def sync(src, dst):
    with torch.no_grad():
        ms, md = src.model.singlegpu_model, dst.model.singlegpu_model
        for a, b in zip(list(ms.parameters()) + list(ms.buffers()), list(md.parameters()) + list(md.buffers())):
            b.copy_(a)
        for os_, od in ((src.optimizer_G, dst.optimizer_G), (src.optimizer_D, dst.optimizer_D)):
            for ps, pd in zip(os_.param_groups[0]["params"], od.param_groups[0]["params"]):
                if ps in os_.state and pd in od.state:
                    for k, v in os_.state[ps].items():
                        od.state[pd][k].copy_(v)


def compare(ta, tb, real, label):
    worst = (0.0, 0.0, None)
    for step in range(16):
        sync(ta, tb)
        a = ta.train_one_step({"real_A": real.clone()}, 0)
        b = tb.train_one_step({"real_A": real.clone()}, 0)
        dl = max(abs(float(a[k]) - float(b[k])) / max(abs(float(a[k])), 1e-2) for k in a)
        group = "Dparams" if step % 2 == 0 else "Gparams"
        ga = torch.cat([p.grad.reshape(-1) for p in getattr(ta, group) if p.grad is not None])
        gb = torch.cat([p.grad.reshape(-1) for p in getattr(tb, group) if p.grad is not None])
        dg = float((ga - gb).norm() / ga.norm().clamp_min(1e-20)) if ga.shape == gb.shape else float("inf")
        if dl > worst[0] or dg > worst[1]:
            worst = (max(dl, worst[0]), max(dg, worst[1]), step)
        if dl > 1e-4 or dg > 1e-3:
            print("   %s step %2d: loss diff %.3e grad diff %.3e  %s" % (label, step, dl, dg, sorted(a)))
    print(" %s: worst loss diff %.3e, worst grad diff %.3e (step %s)" % (label, worst[0], worst[1], worst[2]), flush=True)


def main():
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    real = torch.randn(2, 3, 64, 64, device="cuda", generator=torch.Generator("cuda").manual_seed(5)).clamp(-1, 1)
    t = trainer(False)
    m = t.model.singlegpu_model
    with torch.no_grad():
        for i in range(repeats):
            outs = []
            for _ in range(2):
                sp, gl = m.E(real)
                rec = m.G(sp, gl)
                outs.append((sp, gl, rec, m.D(rec)))
            same = all(torch.equal(x, y) for x, y in zip(*outs))
            print("forward bitwise repeatable (run %d): %s" % (i, same), flush=True)
    for i in range(repeats):
        print("repeat %d" % i)
        compare(trainer(False), trainer(False), real, "eager vs eager")
        tg = trainer(True)
        compare(trainer(False), tg, real, "eager vs graph")
        print(" graphs: captured %s disabled %s" % (sorted(k[0] for k in tg.graphs.captured), tg.graphs.disabled), flush=True)


if __name__ == "__main__":
    main()
