"""CPU: pin the oracle against fixtures produced by the REFERENCE ITSELF (oracle/make_golden.py), fp64."""
import numpy as np
import pytest
import torch

from oracle import sae_oracle as O
from oracle.fixtures import TINY, load_golden, perturbed_state_dict, rel_err, rnd
from swapping_autoencoder_pytorch_b200 import default_options

TOL = 1e-9


def test_fir_torch_and_numpy_against_reference():
    meta, G = load_golden("ops_upfirdn2d")
    for i, c in enumerate(meta["cases"]):
        k = O.make_kernel(c["taps"], torch.float64) * c["gain"]
        x = rnd(meta["x_seed0"] + i, *meta["shape"]).requires_grad_()
        y = O.upfirdn2d(x, k, up=c["up"], down=c["down"], pad=tuple(c["pad"]))
        assert y.shape == G["y%d" % i].shape, c
        assert rel_err(y, G["y%d" % i]) < TOL, c
        w = rnd(meta["w_seed0"] + i, *y.shape)
        gx, = torch.autograd.grad((y * w).sum(), x)
        assert rel_err(gx, G["gx%d" % i]) < TOL, c
        p = c["pad"]
        yn = O.fir_numpy(x.detach().numpy(), k.numpy(), (c["up"],) * 2, (c["down"],) * 2, (p[0], p[1], p[0], p[1]))
        assert rel_err(torch.from_numpy(np.ascontiguousarray(yn)), G["y%d" % i]) < TOL, c


def test_fused_leaky_relu_against_reference():
    meta, G = load_golden("ops_fused_leaky_relu")
    for i, shape in enumerate(meta["shapes"]):
        x = rnd(meta["x_seed0"] + i, *shape).requires_grad_()
        b = rnd(meta["b_seed0"] + i, shape[1]).requires_grad_()
        y = O.fused_leaky_relu(x, b)
        w = rnd(meta["w_seed0"] + i, *shape)
        gx, gb = torch.autograd.grad((y * w).sum(), [x, b])
        assert rel_err(y, G["y%d" % i]) < TOL and rel_err(gx, G["gx%d" % i]) < TOL and rel_err(gb, G["gb%d" % i]) < TOL


def test_layers_against_reference():
    meta, G = load_golden("layers")
    for i, (cin, cout, k, demod, up) in enumerate(meta["modconv"]):
        P = {"m.weight": rnd(400 + i, 1, cout, cin, k, k).requires_grad_(), "m.modulation.weight": rnd(410 + i, cin, 16),
             "m.modulation.bias": rnd(420 + i, cin) * 0.1 + 1}
        x = rnd(430 + i, 2, cin, 6, 7).requires_grad_()
        s = rnd(440 + i, 2, 16).requires_grad_()
        y = O.modulated_conv2d(P, "m", x, s, k, demodulate=demod, upsample=up)
        w = rnd(450 + i, *y.shape)
        gx, gs, gw = torch.autograd.grad((y * w).sum(), [x, s, P["m.weight"]])
        for got, key in ((y, "y"), (gx, "gx"), (gs, "gs"), (gw, "gw")):
            assert rel_err(got, G["modconv%d_%s" % (i, key)]) < TOL, (i, key)
    for i, (cin, cout, blur, refl, down) in enumerate(meta["resblock"]):
        P = {"r.conv1.Conv.weight": rnd(500 + i, cin, cin, 3, 3).requires_grad_(), "r.conv1.Act.bias": rnd(510 + i, cin) * 0.1,
             "r.conv2.Conv.weight": rnd(520 + i, cout, cin, 3, 3), "r.conv2.Act.bias": rnd(530 + i, cout) * 0.1,
             "r.skip.Conv.weight": rnd(540 + i, cout, cin, 1, 1)}
        x = rnd(550 + i, 2, cin, 10, 10).requires_grad_()
        y = O.res_block(P, "r", x, blur_taps=tuple(blur), reflection_pad=refl, downsample=down)
        w = rnd(560 + i, *y.shape)
        gx, = torch.autograd.grad((y * w).sum(), x, create_graph=True)
        gg, = torch.autograd.grad(gx.pow(2).sum(), P["r.conv1.Conv.weight"])
        assert rel_err(y, G["resblock%d_y" % i]) < TOL
        assert rel_err(gx, G["resblock%d_gx" % i]) < TOL
        assert rel_err(gg, G["resblock%d_gg" % i]) < TOL
    for i, up in enumerate([False, True]):
        P = {"s.conv.weight": rnd(600 + i, 1, 8, 8, 3, 3), "s.conv.modulation.weight": rnd(610 + i, 8, 16),
             "s.conv.modulation.bias": torch.ones(8, dtype=torch.float64), "s.noise.weight": torch.tensor([0.3], dtype=torch.float64),
             "s.activate.bias": rnd(620 + i, 8) * 0.1}
        hw = 10 if up else 5
        y = O.styled_conv(P, "s", rnd(630 + i, 2, 8, 5, 5), rnd(640 + i, 2, 16), upsample=up, noise=rnd(650 + i, 2, 1, hw, hw))
        assert rel_err(y, G["styled%d_y" % i]) < TOL
    P = {"l.weight": rnd(700, 8, 16), "l.bias": rnd(701, 8) * 0.1}
    assert rel_err(O.equal_linear(P, "l", rnd(702, 3, 16), activation=True), G["linear_act_y"]) < TOL
    P = {"l.weight": rnd(703, 8, 16), "l.bias": rnd(704, 8)}
    assert rel_err(O.equal_linear(P, "l", rnd(705, 3, 16)), G["linear_y"]) < TOL


def _tiny_model():
    opt = default_options(**TINY)
    sd = perturbed_state_dict(opt)
    return opt, O.OracleModel(opt, sd)


def test_networks_against_reference():
    meta, G = load_golden("networks_tiny")
    opt, m = _tiny_model()
    real = rnd(meta["real_seed"], 2, 3, 64, 64).clamp(-1, 1)
    sp, gl = O.encoder_forward(m.E, opt, real)
    assert rel_err(sp, G["sp"]) < TOL and rel_err(gl, G["gl"]) < TOL
    noises = {}
    for idx, name in enumerate(meta["noise_names"]):        # "<block>.<conv>.noise"
        noises[name[:-len(".noise")]] = rnd(meta["noise_seed0"] + idx, *meta["noise_shapes"][idx])
    rec = O.generator_forward(m.G, opt, sp, gl, noises=noises)
    assert rel_err(rec, G["rec"]) < TOL
    assert rel_err(O.discriminator_forward(m.D, opt, real), G["d_real"]) < TOL
    f1 = O.patch_extract_features(m.Dp, opt, rnd(meta["crop_seeds"][0], 2, 2, 3, 32, 32), aggregate=True)
    f2 = O.patch_extract_features(m.Dp, opt, rnd(meta["crop_seeds"][1], 2, 2, 3, 32, 32))
    assert rel_err(f1, G["patch_feat_agg"]) < TOL and rel_err(f2, G["patch_feat"]) < TOL
    assert rel_err(O.patch_discriminate(m.Dp, f1, f2), G["patch_pred"]) < TOL


def test_loss_graph_against_reference(fp64_default):
    meta, G = load_golden("losses_tiny")
    opt, m = _tiny_model()
    real = rnd(meta["real_seed"], 2, 3, 64, 64).clamp(-1, 1)
    torch.manual_seed(meta["seeds"]["D"])
    for k, v in m.discriminator_losses(real).items():
        assert rel_err(v, G["D/" + k]) < 1e-8, k
    torch.manual_seed(meta["seeds"]["G"])
    for k, v in m.generator_losses(real).items():
        assert rel_err(v, G["G/" + k]) < 1e-8, k
    w = m.D["stylegan2_D.convs.3.conv1.Conv.weight"].requires_grad_()
    wp = m.Dp["convs.2.conv2.Conv.weight"].requires_grad_()
    torch.manual_seed(meta["seeds"]["R1"])
    r1 = m.r1_loss(real)["D_R1"]
    assert rel_err(r1, G["R1/D_R1"]) < 1e-8
    g, gp = torch.autograd.grad(r1.mean(), [w, wp])
    assert rel_err(g, G["R1/grad_D_convs1_conv1"]) < 1e-7
    assert rel_err(gp, G["R1/grad_Dpatch_convs1_conv2"]) < 1e-7
