"""GPU (B200): the kernels either side of the network passes through the C ABI —
``sae_crop_gather`` / ``sae_crop_gather_backward`` against the oracle's F.grid_sample formulation (reference
util/util.py:323-343), ``sae_adam_step`` against ``torch.optim.Adam``, the lazy loss read-back, and the closed-form R1 double
backward of the fused blocks against ordinary autograd through the per-operator nodes."""
import pytest
import torch

from oracle import sae_oracle as O
from oracle.fixtures import TINY, rel_err, rel_l2, rnd
from swapping_autoencoder_pytorch_b200 import backend, default_options, util

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture
def exact_fp32():
    kern = backend.kernels()
    prev, kern.round_tf32 = kern.round_tf32, False
    yield
    kern.round_tf32 = prev


class _Draws:
    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)

    def draw(self, b, lo, hi):
        r = lambda *shape: torch.rand(*shape, generator=self.gen, dtype=torch.float64)      # noqa: E731
        flip = torch.round(r(b, 1, 1, 1)) * 2 - 1.0
        scale = r(b, 1, 1, 2) * (hi - lo) + lo
        offset = (r(b, 1, 1, 2) * 2 - 1) * (1 - scale)
        return flip, scale, offset


@pytest.mark.parametrize("b,res,size,n,strided", [(3, 256, 128, 8, False), (2, 256, 64, 8, True), (2, 64, 32, 2, False),
                                                  (1, 96, 17, 3, True)])
def test_crop_gather_against_oracle(b, res, size, n, strided, exact_fp32, monkeypatch):
    """forward, the zero-padded 32-channel layout, and the adjoint; ``strided``: the source is the 3-channel view of a 4-channel
    channels-last tensor (what the generator's ToRGB hands over)"""
    opt = default_options(patch_size=size, patch_num_crops=n)
    x64 = rnd(5, b, 3, res, res + (8 if strided else 0))
    x64 = x64[..., :res] if strided else x64
    draws = _Draws(11)
    monkeypatch.setattr(O, "draw_crop_parameters", lambda B, o: draws.draw(B, o.patch_min_scale, o.patch_max_scale))
    xr = x64.clone().requires_grad_()
    ref = O.random_crops(xr, opt)
    w = rnd(6, *ref.shape)
    gref, = torch.autograd.grad((ref * w).sum(), xr)
    draws2 = _Draws(11)
    monkeypatch.setattr(util, "draw_crop_parameters",
                        lambda B, sr, device: tuple(t.float().to(device) for t in draws2.draw(B, sr[0], sr[1])))
    if strided:
        base = torch.zeros(b, res, res, 4, device=DEV)
        base[..., :3] = x64.float().to(DEV).permute(0, 2, 3, 1)
        xg = base.permute(0, 3, 1, 2)[:, :3].requires_grad_()
    else:
        xg = x64.float().to(DEV).requires_grad_()
    got = util.apply_random_crop(xg, size, (opt.patch_min_scale, opt.patch_max_scale), num_crops=n)
    assert got.shape == ref.shape
    # fp32 sampling coordinates (as in F.grid_sample's own fp32 path): at 256 pixels one ulp of a coordinate is 3e-5 pixel, and
    # the white-noise test image changes by O(1) per pixel -> 1e-4 of max|ref|; the reference here is fp64 throughout
    tol = 1e-4
    assert rel_err(got, ref) < tol, rel_err(got, ref)
    flat = got.flatten(0, 1)
    from swapping_autoencoder_pytorch_b200.stylegan2_op import conv as C
    assert C._zero_padded_width(flat) == 32
    assert float(flat._base[..., 3:].abs().max()) == 0.0
    g, = torch.autograd.grad((got * w.float().to(DEV)).sum(), xg)
    assert rel_err(g, gref) < tol, rel_err(g, gref)
    # a convolution on the crops takes the padded buffer as it is (no pad kernel) and matches the padded reference
    wt = rnd(7, 32, 3, 3, 3)
    y = C.conv2d(flat, wt.float().to(DEV), padding=1)
    yr = torch.nn.functional.conv2d(ref.detach().flatten(0, 1), wt, padding=1)
    assert rel_err(y, yr) < 2e-3          # TF32 products of 27 white-noise terms


@pytest.mark.parametrize("betas", [(0.0, 0.99), (0.5, 0.9)])
def test_adam_step_against_torch(betas):
    from swapping_autoencoder_pytorch_b200.optimizer import MultiTensorAdam
    shapes = [(64, 32, 3, 3), (7,), (1,), (33, 5), (2048, 130), (3, 1, 1, 1)]
    torch.manual_seed(0)
    ours = [torch.randn(s, device=DEV).requires_grad_() for s in shapes]
    ref = [p.detach().clone().requires_grad_() for p in ours]
    a = MultiTensorAdam(ours, lr=0.002, betas=betas)
    b = torch.optim.Adam(ref, lr=0.002, betas=betas)
    for it in range(6):
        for i, (p, q) in enumerate(zip(ours, ref)):
            g = torch.randn_like(p) * (10.0 ** (i - 3))
            skip = (it % 2 == 1 and i in (1, 3))           # parameters without a gradient keep their own step count
            p.grad = None if skip else g.clone()
            q.grad = None if skip else g.clone()
        a.step()
        b.step()
        for p, q in zip(ours, ref):
            assert rel_err(p, q) < 2e-6, (it, tuple(p.shape), rel_err(p, q))
    sd = a.state_dict()
    ref_sd = b.state_dict()
    for i in ref_sd["state"]:
        assert float(sd["state"][i]["step"]) == float(ref_sd["state"][i]["step"])
        assert rel_err(sd["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"]) < 2e-6
    # bucket-style gradients with the 1/world factor folded in
    views = [torch.full_like(p, 4.0) for p in ours]
    for q in ref:
        q.grad = torch.full_like(q, 2.0)
    a.step(grads=views, grad_scale=0.5)
    b.step()
    for p, q in zip(ours, ref):
        assert rel_err(p, q) < 2e-6
    # the stock optimizer loads this optimizer's state and continues identically
    c = torch.optim.Adam([p.detach().clone().requires_grad_() for p in ours], lr=0.002, betas=betas)
    c.load_state_dict(a.state_dict())
    for p, q in zip(ours, c.param_groups[0]["params"]):
        p.grad = torch.ones_like(p)
        q.grad = torch.ones_like(q)
    a.step()
    c.step()
    for p, q in zip(ours, c.param_groups[0]["params"]):
        assert rel_err(p, q) < 2e-6


def test_lazy_loss_readback():
    losses = {"a": torch.full((4,), 2.0, device=DEV), "b": torch.arange(3.0, device=DEV)}
    out = util.to_numpy(losses, lazy=True)
    assert isinstance(out, util.LazyLosses) and "a" in out and list(out) == ["a", "b"] and len(out) == 2
    assert out._pending is not None                    # nothing waited for yet
    assert float(out["a"]) == 2.0 and float(out["b"]) == 1.0 and out._pending is None
    eager = util.to_numpy(losses)
    assert float(eager["a"]) == 2.0 and float(eager["b"]) == 1.0


def test_closed_form_r1_double_backward_on_gpu():
    """_ResBlockDataGrad (inside data_gradients_only()) against autograd through the per-operator nodes, at a discriminator
    shape (tcgen05 kernels) and a small one: dx, the second-order gradients of all three filters and of the upstream gradient"""
    from swapping_autoencoder_pytorch_b200 import stylegan2_layers as L
    from swapping_autoencoder_pytorch_b200.stylegan2_op import blocks
    import contextlib
    for cin, cout, n, hw in ((128, 256, 2, 64), (32, 64, 3, 33)):
        torch.manual_seed(cin)
        m = L.ResBlock(cin, cout).to(DEV)
        with torch.no_grad():
            m.conv1.Act.bias.normal_(0, 0.1)
            m.conv2.Act.bias.normal_(0, 0.1)
        weights = [m.conv1.Conv.weight, m.conv2.Conv.weight, m.skip.Conv.weight]
        x0 = torch.randn(n, cin, hw, hw, device=DEV)
        w0 = torch.randn(n, cout, hw // 2, hw // 2, device=DEV)
        res = {}
        for mode in ("closed_form", "per_operator"):
            prev = blocks.set_fused_blocks(mode == "closed_form")
            try:
                x, w = x0.clone().requires_grad_(), w0.clone().requires_grad_()
                with (blocks.data_gradients_only() if mode == "closed_form" else contextlib.nullcontext()):
                    gx, = torch.autograd.grad((m(x) * w).sum(), x, create_graph=True)
                    second = torch.autograd.grad(gx.pow(2).sum(), weights + [w])
                res[mode] = [gx.detach()] + list(second)
            finally:
                blocks.set_fused_blocks(prev)
        for i, (a, b) in enumerate(zip(res["closed_form"], res["per_operator"])):
            # same kernels, different TF32 rounding points of intermediates
            assert rel_l2(a, b) < 2e-3 and rel_err(a, b) < 1e-2, (cin, i, rel_l2(a, b), rel_err(a, b))


def test_r1_half_step_time_and_losses_tiny():
    """one D step with R1 through the public driver; the R1 loss agrees between the closed-form and the general path"""
    import swapping_autoencoder_pytorch_b200 as S
    from swapping_autoencoder_pytorch_b200.stylegan2_op import blocks, conv as C
    opt = default_options(**dict(TINY, num_gpus=1, R1_once_every=1))
    torch.manual_seed(0)
    model = S.create_model(opt)
    real = torch.randn(2, 3, 64, 64, device=DEV).clamp(-1, 1)
    inner = model.singlegpu_model
    out = {}
    for mode in ("closed", "general"):
        torch.manual_seed(5)
        for p in inner.parameters():
            p.grad = None
        if mode == "closed":
            r1 = inner(real.clone(), command="compute_R1_loss")["D_R1"]
        else:
            with blocks.per_operator_blocks():
                prev = C.set_data_gradients_only(False)
                try:
                    r1 = inner._compute_R1_loss(real.clone())["D_R1"]
                finally:
                    C.set_data_gradients_only(prev)
        r1.mean().backward()
        out[mode] = (r1.detach().clone(), inner.D.stylegan2_D.convs[1].conv1.Conv.weight.grad.clone(),
                     inner.Dpatch.convs[1].conv2.Conv.weight.grad.clone())
    assert rel_err(out["closed"][0], out["general"][0]) < 2e-3
    assert rel_l2(out["closed"][1], out["general"][1]) < 2e-2 and rel_l2(out["closed"][2], out["general"][2]) < 2e-2


@pytest.mark.parametrize("n,c,hw", [(3, 128, 64), (2, 64, 33), (2, 204, 16), (1, 512, 8), (2, 1024, 5)])
def test_torgb_kernel_against_oracle(n, c, hw, exact_fp32):
    """csrc/torgb.cu: (1) the op with a given per-sample scale — forward and the gradients of input, scale, filter and bias
    against fp64 (plain fp32 FMA arithmetic: 1e-5); (2) through the ToRGB module, whose style path adds a TF32 linear — against
    the oracle's modulated_conv2d(demodulate=False) + bias at the per-op TF32 tolerance"""
    from swapping_autoencoder_pytorch_b200 import stylegan2_layers as L
    from swapping_autoencoder_pytorch_b200.stylegan2_op import torgb
    x, s, w, b = rnd(5, n, c, hw, hw + 1), rnd(6, n, c) * 0.3 + 1, rnd(1, 3, c, 1, 1), rnd(4, 1, 3, 1, 1) * 0.1
    wscale = 1.0 / c ** 0.5
    ref_in = [t.clone().requires_grad_() for t in (x, s, w, b)]
    y_ref = torch.nn.functional.conv2d(ref_in[0] * ref_in[1][:, :, None, None], ref_in[2] * wscale) + ref_in[3]
    wgt = rnd(7, *y_ref.shape)
    gref = torch.autograd.grad((y_ref * wgt).sum(), ref_in)
    got_in = [t.float().to(DEV).requires_grad_() for t in (x, s, w, b)]
    y = torgb(*got_in, wscale)
    assert y.shape == y_ref.shape and rel_err(y, y_ref) < 1e-5, rel_err(y, y_ref)
    got = torch.autograd.grad((y * wgt.float().to(DEV)).sum(), got_in)
    for name, a, r in zip(("x", "scale", "weight", "bias"), got, gref):
        assert rel_err(a, r) < 2e-5, (name, rel_err(a, r))
    # the gradient may arrive in any layout (crop adjoint: NCHW-contiguous; image losses: channels-last views)
    g2, = torch.autograd.grad(torgb(*got_in, wscale), got_in[0], wgt.float().to(DEV).contiguous())
    assert rel_err(g2, gref[0]) < 2e-5
    # (2) module level
    P = {"conv.weight": w[None], "conv.modulation.weight": rnd(2, c, 16), "conv.modulation.bias": rnd(3, c) * 0.1 + 1, "bias": b}
    m = L.ToRGB(c, 16, upsample=False)
    sd = m.state_dict()
    sd.update({k: v.float() for k, v in P.items()})
    m.load_state_dict(sd)
    m = m.to(DEV)
    st = rnd(8, n, 16)
    xr, sr = x.clone().requires_grad_(), st.clone().requires_grad_()
    y_ref = O.modulated_conv2d({"t." + k: v for k, v in P.items()}, "t.conv", xr, sr, 1, demodulate=False) + P["bias"]
    gxr, gsr = torch.autograd.grad((y_ref * wgt).sum(), [xr, sr])
    xg, sg = x.float().to(DEV).requires_grad_(), st.float().to(DEV).requires_grad_()
    ym = m(xg, sg)
    gx, gs = torch.autograd.grad((ym * wgt.float().to(DEV)).sum(), [xg, sg])
    assert rel_err(ym, y_ref) < 1e-3 and rel_err(gx, gxr) < 2e-3 and rel_err(gs, gsr) < 3e-3, (rel_err(ym, y_ref), rel_err(gx, gxr), rel_err(gs, gsr))


@pytest.mark.parametrize("kh,c,h,pad,with_noise", [(4, 128, 65, (1, 1), True), (4, 32, 33, (1, 1), False), (3, 64, 40, (1, 1), True),
                                                   (4, 512, 17, (1, 1), True)])
def test_fir_bias_act_fused(kh, c, h, pad, with_noise, exact_fp32):
    """sae_fir_bias_act == sae_upfirdn2d_separable followed by sae_fused_bias_act (noise + bias + leaky-ReLU), and the
    autograd Function around it (forward, input / bias / noise-weight gradients) against the oracle in fp64"""
    from swapping_autoencoder_pytorch_b200.stylegan2_op.blocks import FirSpec, fir_noise_bias_act
    kern = backend.kernels()
    taps1 = [1.0, 3.0, 3.0, 1.0] if kh == 4 else [1.0, 2.0, 1.0]
    t = tuple(2.0 * v / sum(taps1) for v in taps1)                       # the generator's blur carries the x4 upsampling gain
    kernel = torch.outer(torch.tensor(t), torch.tensor(t)).to(DEV)
    n = 3
    x = torch.randn(n, h, h + 2, c, device=DEV)
    oh, ow = h + pad[0] + pad[1] - kh + 1, h + 2 + pad[0] + pad[1] - kh + 1
    bias = torch.randn(c, device=DEV) * 0.1
    noise = torch.randn(n, oh, ow, device=DEV) if with_noise else None
    nw = torch.tensor([0.3], device=DEV) if with_noise else None
    fused = kern.fir_bias_act(x, (t, t), (pad[0], pad[1], pad[0], pad[1]), bias, noise.reshape(-1) if with_noise else None, nw, 0.2, 1.3)
    assert fused is not None
    y = kern.upfirdn2d(x, kernel, 1, 1, 1, 1, pad[0], pad[1], pad[0], pad[1], taps=(t, t))
    ref = kern.bias_act(y, bias, None, 3, 0, 0.2, 1.3, noise=noise.reshape(-1) if with_noise else None, noise_weight=nw)
    assert rel_err(fused, ref) < 2e-6, rel_err(fused, ref)
    # autograd Function against the fp64 oracle
    xr = x.double().cpu().permute(0, 3, 1, 2).requires_grad_()
    br = bias.double().cpu().requires_grad_()
    nwr = nw.double().cpu().requires_grad_() if with_noise else None
    yr = O.upfirdn2d(xr, kernel.double().cpu(), pad=pad)
    if with_noise:
        yr = yr + nwr * noise.double().cpu().unsqueeze(1)
    yr = O.fused_leaky_relu(yr, br, 0.2, 1.3)
    wgt = torch.randn_like(yr)
    gref = torch.autograd.grad((yr * wgt).sum(), [xr, br] + ([nwr] if with_noise else []))
    xg = x.permute(0, 3, 1, 2).requires_grad_()
    bg = bias.clone().requires_grad_()
    nwg = nw.clone().requires_grad_() if with_noise else None
    out = fir_noise_bias_act(xg, FirSpec(kernel, pad, (t, t), 1), noise.unsqueeze(1) if with_noise else None, nwg, bg, 0.2, 1.3)
    assert rel_err(out, yr) < 2e-6
    got = torch.autograd.grad((out * wgt.float().to(DEV)).sum(), [xg, bg] + ([nwg] if with_noise else []))
    for a, b in zip(got, gref):
        assert rel_err(a, b) < 2e-5, rel_err(a, b)


@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 128, 128, 128, 128), (3, 64, 128, 64, 128), (4, 256, 256, 128, 128), (2, 64, 64, 64, 64)])
def test_modulated_conv_per_sample_filters(n, cin, cout, h, w):
    """sae_filter_modulate + sae_conv2d_fprop_per_sample / dgrad_per_sample / wgrad_modulated.  (1) ModulatedConv2d alone — a
    bilinear map, so every gradient is held to the per-op TF32 tolerance against the oracle in fp64: output, dx, dstyle, dweight,
    d(modulation weights).  Batch sizes and maps are chosen so that weight-gradient CTAs cross image boundaries (several
    accumulator drains per CTA).  (2) Through StyledConv (noise + bias + leaky-ReLU in the same kernel's epilogue): output at
    1e-3; gradients in relative L2 only — a TF32-level difference in a pre-activation near zero flips that element's mask."""
    from swapping_autoencoder_pytorch_b200 import stylegan2_layers as L
    P = {"conv.weight": rnd(1, 1, cout, cin, 3, 3), "conv.modulation.weight": rnd(2, cin, 16), "conv.modulation.bias": rnd(3, cin) * 0.1 + 1,
         "noise.weight": torch.tensor([0.3], dtype=torch.float64), "activate.bias": rnd(4, cout) * 0.1}
    x, st, nz = rnd(5, n, cin, h, w), rnd(6, n, 16), rnd(7, n, 1, h, w)
    # (1) bare modulated convolution
    mc = L.ModulatedConv2d(cin, cout, 3, 16)
    sd = mc.state_dict()
    sd.update({k[len("conv."):]: v.float() for k, v in P.items() if k.startswith("conv.")})
    mc.load_state_dict(sd)
    mc = mc.to(DEV)
    xg, sg = x.float().to(DEV).requires_grad_(), st.float().to(DEV).requires_grad_()
    assert mc.per_sample_geom(xg, sg) is not None, "shape should take the per-sample-filter path"
    Pr = {k: v.clone().requires_grad_() for k, v in P.items()}
    xr, sr = x.clone().requires_grad_(), st.clone().requires_grad_()
    y_ref = O.modulated_conv2d({"t." + k: v for k, v in Pr.items()}, "t.conv", xr, sr, 3)
    wgt = rnd(8, *y_ref.shape)
    ref = torch.autograd.grad((y_ref * wgt).sum(), [xr, sr, Pr["conv.weight"], Pr["conv.modulation.weight"]])
    y = mc(xg, sg)
    got = torch.autograd.grad((y * wgt.float().to(DEV)).sum(), [xg, sg, mc.weight, mc.modulation.weight])
    assert rel_err(y, y_ref) < 1e-3, rel_err(y, y_ref)
    for name, a, b, tol in zip(("x", "style", "weight", "modulation"), got, ref, (2e-3, 3e-3, 2e-3, 3e-3)):
        assert rel_err(a, b) < tol, (name, rel_err(a, b), rel_l2(a, b))
    # (2) with the StyledConv tail in the epilogue
    m = L.StyledConv(cin, cout, 3, 16)
    sd = m.state_dict()
    sd.update({k: v.float() for k, v in P.items()})
    m.load_state_dict(sd)
    m = m.to(DEV)
    Pr = {k: v.clone().requires_grad_() for k, v in P.items()}
    xr, sr = x.clone().requires_grad_(), st.clone().requires_grad_()
    y_ref = O.styled_conv({"t." + k: v for k, v in Pr.items()}, "t", xr, sr, noise=nz)
    names = ["conv.weight", "conv.modulation.weight", "noise.weight", "activate.bias"]
    ref = torch.autograd.grad((y_ref * wgt).sum(), [xr, sr] + [Pr[k] for k in names])
    xg, sg = x.float().to(DEV).requires_grad_(), st.float().to(DEV).requires_grad_()
    y = m(xg, sg, noise=nz.float().to(DEV))
    got = torch.autograd.grad((y * wgt.float().to(DEV)).sum(), [xg, sg, m.conv.weight, m.conv.modulation.weight, m.noise.weight, m.activate.bias])
    assert rel_err(y, y_ref) < 1e-3, rel_err(y, y_ref)
    for name, a, b in zip(["x", "style"] + names, got, ref):
        if name == "noise.weight":
            # one scalar = a sum of n*h*w*cout signed terms of unit size: judged against that sum's natural scale
            assert abs(float(a) - float(b)) < 5e-2 * (y_ref.numel() ** 0.5), (name, float(a), float(b))
            continue
        assert rel_l2(a, b) < 2e-2, (name, rel_l2(a, b), rel_err(a, b))


def _unpack_mask(mask, shape):
    """int32 words -> bool tensor of ``shape`` (bit i & 31 of word i >> 5 = element i)"""
    bits = (mask.view(-1, 1) >> torch.arange(32, device=mask.device, dtype=torch.int32).view(1, 32)) & 1
    return bits.reshape(shape).bool()


@pytest.mark.parametrize("n,h,c,k,r,stride,pad,per_sample", [
    (4, 64, 128, 128, 3, 1, 1, False),        # shared-window pair kernel (conv_tc5<128>)
    (3, 40, 64, 64, 3, 1, 1, False),          # conv_tc5<64>, ragged 16 x 8 tiling
    (6, 128, 32, 32, 3, 1, 1, False),         # conv_tc5<32> with the resident filter slice
    (4, 65, 128, 256, 3, 2, 0, False),        # stride 2: conv_tc6<256>, two epilogue groups
    (4, 33, 64, 128, 3, 2, 0, False),         # stride 2: conv_tc6<128>
    (2, 6, 64, 96, 3, 1, 1, False),           # small map: one-tile kernel, 96 output channels
    (5, 64, 32, 128, 1, 1, 0, False),         # 1x1 (FromRGB shape)
    (2, 64, 128, 128, 3, 1, 1, True),         # per-sample filters (modulated convolution)
])
def test_activation_bit_mask_matches_output_sign(n, h, c, k, r, stride, pad, per_sample):
    """sae_conv_epilogue.act_mask: every tensor-core conv that applies the leaky-ReLU also writes 1 bit per output element
    (positive branch or not); sae_bias_act_backward / sae_fir_act_backward read it instead of the 4-byte output.  The mask must
    equal ``out > 0`` everywhere, and the masked-gradient pass must give the same bytes either way."""
    kern = backend.kernels()
    if not kern.act_masks:
        pytest.skip("activation bit masks switched off (SAE_ACT_MASK=0)")
    g = backend.make_geom(n, h, h, c, k, r, r, stride, pad, pad)
    x = rnd(1, n, h, h, c).float().to(DEV)
    bias = (rnd(2, k) * 0.3).float().to(DEV)
    noise = rnd(3, n * g.P * g.Q).float().to(DEV)
    nw = torch.tensor([0.4], device=DEV)
    if per_sample:
        w = (rnd(4, n, k, r, r, c) / (c * r * r) ** 0.5).float().to(DEV)
        y = kern.conv_fprop_per_sample(x, w, g, bias=bias, act=3, alpha=0.2, gain=1.4, noise=noise, noise_weight=nw)
    else:
        w = (rnd(4, k, r, r, c) / (c * r * r) ** 0.5).float().to(DEV)
        y = kern.conv_fprop(x, w, g, bias=bias, act=3, alpha=0.2, gain=1.4, noise=noise, noise_weight=nw)
    mask = backend.act_mask_of(y)
    assert mask is not None and mask.numel() * 32 == y.numel()
    assert torch.equal(_unpack_mask(mask, y.shape), y > 0)
    dy = rnd(5, *y.shape).float().to(DEV)
    a = kern.bias_act_backward(dy, y, 0.2, 1.4, want_bias=True, noise=noise, mask=mask)
    b = kern.bias_act_backward(dy, y, 0.2, 1.4, want_bias=True, noise=noise)
    assert torch.equal(a[0], b[0])
    assert rel_err(a[1], b[1]) < 1e-5 and rel_err(a[2], b[2]) < 1e-5          # (atomics: summation order differs run to run)
    # linear epilogue: no mask
    assert backend.act_mask_of(kern.conv_fprop(x, w[0] if per_sample else w, g, bias=bias)) is None


def test_activation_bit_mask_through_the_fir_kernels():
    """sae_fir_bias_act writes the mask, sae_fir_act_backward reads it (the blur adjoint in front of a ResBlock's first activation)"""
    from swapping_autoencoder_pytorch_b200.stylegan2_op.upfirdn2d import _flip_taps
    kern = backend.kernels()
    if not kern.act_masks:
        pytest.skip("activation bit masks switched off (SAE_ACT_MASK=0)")
    t = [1.0, 3.0, 3.0, 1.0]
    taps = ([v / 8 for v in t], [v / 8 for v in t])
    x = rnd(1, 3, 37, 41, 64).float().to(DEV)
    bias = (rnd(2, 64) * 0.3).float().to(DEV)
    noise = rnd(3, 3 * 36 * 40).float().to(DEV)
    out = kern.fir_bias_act(x, taps, (1, 1, 1, 1), bias, noise, torch.tensor([0.5], device=DEV), 0.2, 1.4)
    assert tuple(out.shape) == (3, 36, 40, 64)
    mask = backend.act_mask_of(out)
    assert mask is not None and torch.equal(_unpack_mask(mask, out.shape), out > 0)
    # the blur adjoint of a gradient of the blurred shape lands on the activation's shape
    g = rnd(4, 3, 36 + 3 - 2 - 2 + 0, 40 + 3 - 2 - 2 + 0, 64).float().to(DEV)       # [3, 35, 39, 64] -> pad (2, 2) -> 36 x 40
    a = kern.fir_act_backward(g, _flip_taps(taps), out, (2, 2, 2, 2), 0.2, 1.4, want_bias=True, mask=mask)
    b = kern.fir_act_backward(g, _flip_taps(taps), out, (2, 2, 2, 2), 0.2, 1.4, want_bias=True)
    assert a is not None and b is not None
    assert torch.equal(a[0], b[0]) and rel_err(a[1], b[1]) < 1e-5
