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


def test_graph_replay_matches_eager_on_a_deterministic_step(monkeypatch):
    """Without crops (no patch discriminator) and with the noise maps zeroed the half-steps draw no random numbers, so
    eager and captured execution must follow the same loss trajectory (Adam at beta1 = 0 is sign-like on tiny
    gradients, so the comparison is on losses, not on individual parameters)."""
    from swapping_autoencoder_pytorch_b200.stylegan2_layers import NoiseInjection

    def zero_noise(self, image, noise=None):
        if self.image_size is None:
            self.image_size = image.shape
        b, _, h, w = image.shape
        return image.new_empty(b, 1, h, w).zero_()
    monkeypatch.setattr(NoiseInjection, "resolve_noise", zero_noise)
    det = dict(lambda_PatchGAN=0.0, lambda_patch_R1=0.0, R1_once_every=2)
    real = torch.randn(2, 3, 64, 64, device=DEV, generator=torch.Generator(DEV).manual_seed(5)).clamp(-1, 1)
    runs = {}
    for mode in (False, True):
        tr = _trainer(cuda_graphs=mode, **det)
        hist = []
        for _ in range(14):
            hist.append(tr.train_one_step({"real_A": real.clone()}, 0))
        runs[mode] = (tr, hist)
    tr, hist = runs[True]
    assert tr.graphs is not None and tr.graphs.disabled is None, tr.graphs and tr.graphs.disabled
    assert {k[0] for k in tr.graphs.captured} == {"D", "G", "R1"}, list(tr.graphs.captured)
    assert tr.graphs.replayed_launches > 0
    for step, (a, b) in enumerate(zip(runs[False][1], hist)):
        assert a.keys() == b.keys(), (step, a.keys(), b.keys())
        for k in a:
            fa, fb = float(a[k]), float(b[k])
            assert math.isfinite(fb), (step, k, fb)
            assert abs(fa - fb) <= 2e-2 * max(abs(fa), 1e-2), (step, k, fa, fb)


def test_graph_replay_full_model_with_crops_and_noise():
    """The complete loss graph (crops, noise, patch discriminator, R1 on both discriminators) under replay: fresh random
    draws every step, finite losses, parameters moving."""
    tr = _trainer(cuda_graphs=True, R1_once_every=2)
    real = torch.randn(2, 3, 64, 64, device=DEV).clamp(-1, 1)
    tr.graphs.warm_up(real)
    assert tr.graphs.disabled is None, tr.graphs.disabled
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
