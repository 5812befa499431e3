"""GPU (B200): CUDA-graph execution of the half-steps (graphs.py) against eager execution of the same bodies."""
import math

import pytest
import torch

from oracle.fixtures import TINY
from swapping_autoencoder_pytorch_b200 import default_options

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _trainer(**over):
    import swapping_autoencoder_pytorch_b200 as S
    opt = default_options(**dict(TINY, num_gpus=1, **over))
    torch.manual_seed(0)
    model = S.create_model(opt)
    return S.create_optimizer(opt, model)


def _sync_state(src, dst):
    """parameters, buffers and Adam state of trainer ``src`` copied into trainer ``dst`` (same architecture)"""
    with torch.no_grad():
        ms, md = src.model.singlegpu_model, dst.model.singlegpu_model
        for a, b in zip(list(ms.parameters()) + list(ms.buffers()), list(md.parameters()) + list(md.buffers())):
            b.copy_(a)
        for os_, od in ((src.optimizer_G, dst.optimizer_G), (src.optimizer_D, dst.optimizer_D)):
            a, b = os_._state(), od._state()           # flat moments + per-parameter step counts (optimizer.MultiTensorAdam)
            b.exp_avg.copy_(a.exp_avg)
            b.exp_avg_sq.copy_(a.exp_avg_sq)
            b.steps.copy_(a.steps)


def _teacher_forced_run(real, det):
    """16 half-steps of an eager and a graph trainer, the graph trainer re-seeded with the eager one's state before each;
    returns (graph trainer, [(step, kind, how, worst relative loss difference, gradient rel-L2 difference)])"""
    te, tg = _trainer(cuda_graphs=False, **det), _trainer(cuda_graphs=True, **det)
    rows = []
    for step in range(16):
        _sync_state(te, tg)
        captured_before = set(k[0] for k in tg.graphs.captured)
        a = te.train_one_step({"real_A": real.clone()}, 0)
        b = tg.train_one_step({"real_A": real.clone()}, 0)
        assert a.keys() == b.keys(), (step, a.keys(), b.keys())
        dl = 0.0
        for k in a:
            fa, fb = float(a[k]), float(b[k])
            assert math.isfinite(fb), (step, k, fb)
            dl = max(dl, abs(fa - fb) / max(abs(fa), 1e-2))
        group = "Dparams" if step % 2 == 0 else "Gparams"
        ga = torch.cat([p.grad.reshape(-1) for p in getattr(te, group) if p.grad is not None])
        gb = torch.cat([p.grad.reshape(-1) for p in getattr(tg, group) if p.grad is not None])
        assert ga.shape == gb.shape, (step, ga.shape, gb.shape)
        dg = float((ga - gb).norm() / ga.norm().clamp_min(1e-20))
        kind = ("D+R1" if "D_R1" in a else "D") if step % 2 == 0 else "G"
        how = "replay" if kind[0] in captured_before else ("capture" if kind[0] in set(k[0] for k in tg.graphs.captured) else "eager")
        rows.append((step, kind, how, dl, dg))
    return tg, rows


def test_graph_replay_matches_eager_step_by_step(monkeypatch):
    """Teacher-forced comparison: before every half-step the graph trainer receives the eager trainer's parameters and
    Adam state, then both run the step on the same images.  Without crops (no patch discriminator) and with the noise
    maps zeroed the step draws no random numbers, so losses and gradients must agree to kernel-level noise (fp32 atomics
    in the weight-gradient reductions).  Trajectories are NOT compared: Adam at beta1 = 0 is sign-like on tiny gradients
    and two eager runs drift apart just the same."""
    from swapping_autoencoder_pytorch_b200.stylegan2_layers import NoiseInjection

    def zero_noise(self, image, noise=None):
        if self.image_size is None:
            self.image_size = image.shape
        b, _, h, w = image.shape
        return image.new_empty(b, 1, h, w).zero_()
    monkeypatch.setattr(NoiseInjection, "resolve_noise", zero_noise)
    det = dict(lambda_PatchGAN=0.0, lambda_patch_R1=0.0, R1_once_every=2)
    real = torch.randn(2, 3, 64, 64, device=DEV, generator=torch.Generator(DEV).manual_seed(5)).clamp(-1, 1)

    tg, rows = _teacher_forced_run(real, det)
    assert tg.graphs.disabled is None, (tg.graphs.disabled, tg.graphs.last_traceback)
    assert {k[0] for k in tg.graphs.captured} == {"D", "G", "R1"}, sorted(tg.graphs.captured)
    assert tg.graphs.replayed_launches > 0
    # Bounds.  Plain D and G half-steps: kernel noise only (measured < 1e-4 / 1e-3).  A D half-step WITH lazy R1 evaluates
    # the R1 loss on the parameters the D update of the same half-step has just produced: that Adam update (beta1 = 0) is
    # sign-like on tiny gradients, so two EAGER runs from identical state already differ there by up to 9e-4 in the loss and
    # 5e-3 in the gradient (scripts/determinism_probe.py on the B200) — the R1 rows get the correspondingly wider bound.
    bad = [r for r in rows if not ((r[3] < 1e-2 and r[4] < 5e-2) if r[1] == "D+R1" else (r[3] < 1e-3 and r[4] < 1e-2))]
    assert not bad, (bad, rows)


def test_graph_replay_full_model_with_crops_and_noise():
    """The complete loss graph (crops, noise, patch discriminator, R1 on both discriminators) under replay: fresh random
    draws every step, finite losses, parameters moving."""
    tr = _trainer(cuda_graphs=True, R1_once_every=2)
    real = torch.randn(2, 3, 64, 64, device=DEV).clamp(-1, 1)
    tr.graphs.warm_up(real)
    assert tr.graphs.disabled is None, (tr.graphs.disabled, tr.graphs.last_traceback)
    assert len(tr.graphs.captured) == 3
    w = tr.Dparams[0].detach().clone()
    seen = []
    for _ in range(6):
        out = tr.train_one_step({"real_A": real}, 0)
        assert all(math.isfinite(float(v)) for v in out.values()), out
        if "PatchD_mix" in out:
            seen.append(float(out["PatchD_mix"]))
    assert not torch.equal(w, tr.Dparams[0].detach())
    assert len(set(seen)) == len(seen), seen        # new crops / noise every replay -> the loss never repeats exactly
