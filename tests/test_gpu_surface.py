"""GPU (B200): (1) the public drop-in ops in the mode that SHIPS — ``round_tf32=True``, every op output rounded to the nearest
TF32 value — against the fp64 oracle with the rounding's own bound 2^-11 (+ fp32 slack);
(2) SURVEY.md §8 f4: the operator-surface branches training does not exercise (Upsample / Downsample, ToRGB with an upsampled
skip, spatially varying style, downsampling ModulatedConv2d, the original StyleGAN2 ``Generator``, encoder feature
extraction) THROUGH THE CUDA KERNELS against the reference-generated goldens (``tests/golden/surface_branches.npz``,
``encoder_features_tiny.npz``; the CPU suite checks the same goldens on the kernel emulation at 1e-9)."""
import numpy as np
import pytest
import torch

from oracle import sae_oracle as O
from oracle.fixtures import load_golden, perturbed_state_dict, rel_err, rnd
from swapping_autoencoder_pytorch_b200 import backend, default_options

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_ROUNDED = 5e-4        # one round-to-nearest to TF32 (10-bit mantissa): 2^-11 = 4.88e-4 relative, plus fp32 arithmetic
TOL_TF32 = 1e-3
TOL_NET = 3e-3


def cuda(t):
    return t.float().to(DEV)


def _load(module, params):
    sd = module.state_dict()
    sd.update({k: v.float() for k, v in params.items()})
    module.load_state_dict(sd)
    return module.to(DEV)


# ------------------------------------------------------------------------------------------------ (1) default mode
def test_default_mode_is_rounding():
    assert backend.kernels().round_tf32 is True


@pytest.mark.parametrize("taps,pad,c,h,up,down", [([1, 3, 3, 1], (2, 2), 128, 64, 1, 1), ([1, 3, 3, 1], (1, 1), 64, 65, 1, 1),
                                                  ([1, 2, 1], (1, 0), 256, 16, 1, 1), ([1, 3, 3, 1], (1, 1), 3, 33, 1, 1),
                                                  ([1, 3, 3, 1], (1, 1), 32, 63, 1, 2), ([1, 3, 3, 1], (2, 1), 8, 9, 2, 1)])
def test_upfirdn2d_default_mode(taps, pad, c, h, up, down):
    from swapping_autoencoder_pytorch_b200.stylegan2_op import upfirdn2d
    k = O.make_kernel(taps, torch.float64) * (up ** 2)
    x = rnd(5, 2, c, h, h + 1)
    xr = x.clone().requires_grad_()
    y_ref = O.upfirdn2d(xr, k, up=up, down=down, pad=pad)
    w = rnd(6, *y_ref.shape)
    g_ref, = torch.autograd.grad((y_ref * w).sum(), xr)
    xg = cuda(x).requires_grad_()
    t1 = tuple(v / sum(taps) * up for v in taps)
    for kw in ({}, {"taps": (t1, t1)}):
        y = upfirdn2d(xg, cuda(k), up=up, down=down, pad=pad, **kw)
        assert y.shape == y_ref.shape and rel_err(y, y_ref) < TOL_ROUNDED, (kw, rel_err(y, y_ref))
        g, = torch.autograd.grad((y * cuda(w)).sum(), xg)
        assert rel_err(g, g_ref) < TOL_ROUNDED, (kw, rel_err(g, g_ref))
        # the stored values are exactly TF32-representable (low 13 mantissa bits zero)
        assert int((y.detach().contiguous().view(torch.int32) & 0x1FFF).abs().max()) == 0


@pytest.mark.parametrize("shape", [(2, 4, 5, 6), (3, 8), (4, 128, 32, 32), (16, 2048)])
def test_fused_leaky_relu_default_mode(shape):
    from swapping_autoencoder_pytorch_b200.stylegan2_op import fused_leaky_relu
    x, b, w = rnd(1, *shape), rnd(2, shape[1]), rnd(3, *shape)
    xr, br = x.clone().requires_grad_(), b.clone().requires_grad_()
    y_ref = O.fused_leaky_relu(xr, br)
    gx_r, gb_r = torch.autograd.grad((y_ref * w).sum(), [xr, br])
    xg, bg = cuda(x).requires_grad_(), cuda(b).requires_grad_()
    y = fused_leaky_relu(xg, bg)
    assert rel_err(y, y_ref) < TOL_ROUNDED
    gx, gb = torch.autograd.grad((y * cuda(w)).sum(), [xg, bg])
    assert rel_err(gx, gx_r) < TOL_ROUNDED and rel_err(gb, gb_r) < TOL_ROUNDED


def test_modulated_conv_default_mode():
    """x * s then conv: default rounding mode, per-op TF32 tolerance"""
    from swapping_autoencoder_pytorch_b200 import stylegan2_layers as L
    for cin, cout, hw in ((32, 64, 24), (128, 128, 32)):
        P = {"weight": rnd(1, 1, cout, cin, 3, 3), "modulation.weight": rnd(2, cin, 16), "modulation.bias": rnd(3, cin) * 0.1 + 1}
        m = _load(L.ModulatedConv2d(cin, cout, 3, 16), P)
        x, s = rnd(4, 3, cin, hw, hw), rnd(5, 3, 16)
        xr, sr = x.clone().requires_grad_(), s.clone().requires_grad_()
        y_ref = O.modulated_conv2d({"m." + k: v for k, v in P.items()}, "m", xr, sr, 3)
        w = rnd(6, *y_ref.shape)
        gx_r, gs_r = torch.autograd.grad((y_ref * w).sum(), [xr, sr])
        xg, sg = cuda(x).requires_grad_(), cuda(s).requires_grad_()
        y = m(xg, sg)
        gx, gs = torch.autograd.grad((y * cuda(w)).sum(), [xg, sg])
        assert rel_err(y, y_ref) < TOL_TF32 and rel_err(gx, gx_r) < 2 * TOL_TF32 and rel_err(gs, gs_r) < 3 * TOL_TF32, \
            (cin, rel_err(y, y_ref), rel_err(gx, gx_r), rel_err(gs, gs_r))


# ------------------------------------------------------------------------------------------------ (2) f4 branches
def test_surface_branches_on_gpu():
    from swapping_autoencoder_pytorch_b200 import stylegan2_layers as L
    meta, G = load_golden("surface_branches")
    x = cuda(rnd(1000, 2, 4, 7, 6)).requires_grad_()
    for name, mod in (("upsample", L.Upsample([1, 3, 3, 1])), ("downsample", L.Downsample([1, 3, 3, 1]))):
        mod = mod.to(DEV)
        y = mod(x)
        gx, = torch.autograd.grad((y * cuda(rnd(1001, *y.shape))).sum(), x)
        assert rel_err(y, G[name + "_y"]) < TOL_ROUNDED and rel_err(gx, G[name + "_gx"]) < TOL_ROUNDED, name
    m = _load(L.ToRGB(8, 16, upsample=True), {"conv.weight": rnd(1010, 1, 3, 8, 1, 1), "conv.modulation.weight": rnd(1011, 8, 16),
                                              "conv.modulation.bias": rnd(1012, 8) * 0.1 + 1, "bias": rnd(1013, 1, 3, 1, 1) * 0.1})
    y = m(cuda(rnd(1014, 2, 8, 10, 10)), cuda(rnd(1015, 2, 16)), skip=cuda(rnd(1016, 2, 3, 5, 5)))
    assert rel_err(y, G["torgb_skip_y"]) < TOL_TF32, rel_err(y, G["torgb_skip_y"])
    P = {"weight": rnd(1020, 1, 12, 8, 3, 3), "modulation.weight": rnd(1021, 8, 16), "modulation.bias": rnd(1022, 8) * 0.1 + 1}
    m = _load(L.ModulatedConv2d(8, 12, 3, 16), P)
    xs = cuda(rnd(1023, 1, 8, 6, 7)).requires_grad_()
    ss = cuda(rnd(1024, 1, 16, 3, 4)).requires_grad_()
    y = m(xs, ss)
    gx, gs = torch.autograd.grad((y * cuda(rnd(1025, *y.shape))).sum(), [xs, ss])
    assert rel_err(y, G["modconv_spatial_y"]) < TOL_TF32
    assert rel_err(gx, G["modconv_spatial_gx"]) < 2 * TOL_TF32 and rel_err(gs, G["modconv_spatial_gs"]) < 3 * TOL_TF32
    y2 = m(torch.cat([xs, xs]), torch.cat([ss, ss]))
    assert rel_err(y2[1], y[0]) < 1e-5
    m = _load(L.ModulatedConv2d(8, 12, 3, 16, downsample=True), P)
    y = m(cuda(rnd(1026, 2, 8, 8, 8)), cuda(rnd(1027, 2, 16)))
    assert rel_err(y, G["modconv_down_y"]) < TOL_TF32
    g = L.Generator(8, 16, 2, channel_multiplier=1)
    sd = g.state_dict()
    keys = sorted(k for k in sd if sd[k].dtype.is_floating_point and not k.endswith(".kernel"))
    assert keys == meta["keys"]
    rs = np.random.RandomState(meta["state_seed"])
    for k in keys:
        sd[k] = torch.from_numpy(rs.standard_normal(tuple(sd[k].shape))).float() * (0.1 if k.endswith("bias") or "noise" in k else 1.0)
    g.load_state_dict(sd)
    g = g.to(DEV)
    img, _ = g([cuda(rnd(1031, 2, 16))], randomize_noise=False)
    assert rel_err(img, G["generator8_img"]) < TOL_NET, rel_err(img, G["generator8_img"])


def test_encoder_feature_extraction_on_gpu():
    import swapping_autoencoder_pytorch_b200 as S
    meta, G = load_golden("encoder_features_tiny")
    opt = default_options(**dict(meta["opt"], num_gpus=1))
    model = S.create_model(opt).singlegpu_model
    sd = perturbed_state_dict(opt, dtype=torch.float64, param_seed=meta["param_seed"], bias_seed=meta["bias_seed"])
    own = model.state_dict()
    own.update({k: v.float().to(DEV) for k, v in sd.items() if k in own})
    model.load_state_dict(own)
    real = cuda(rnd(meta["real_seed"], 2, 3, 64, 64).clamp(-1, 1))
    with torch.no_grad():
        sp, gl, feat = model.encode(real, extract_features=True)
    assert rel_err(sp, G["sp"]) < TOL_NET and rel_err(gl, G["gl"]) < TOL_NET and rel_err(feat, G["feature"]) < TOL_NET


def test_inference_callers_on_gpu():
    """encode / decode / fix_noise / get_visuals_for_snapshot of the model (reference swapping_autoencoder_model.py:233-264)"""
    import swapping_autoencoder_pytorch_b200 as S
    from oracle.fixtures import TINY
    opt = default_options(**dict(TINY, num_gpus=1, isTrain=True))
    torch.manual_seed(0)
    model = S.create_model(opt)
    real = torch.randn(4, 3, 64, 64, device=DEV).clamp(-1, 1)
    with torch.no_grad():
        sp, gl = model(real, command="encode")
        params = model(real, command="fix_noise")
        assert len(params) > 0 and all(p.is_cuda for p in params)
        a = model(sp, gl, command="decode")
        b = model(sp, gl, command="decode")
        assert torch.equal(a, b)                     # fixed noise: decoding is deterministic
        vis = model(real, command="get_visuals_for_snapshot")
    assert set(vis) >= {"real", "rec", "mix"} and vis["rec"].shape == real.shape and vis["mix"].shape == real.shape
    assert torch.isfinite(vis["mix"]).all()
