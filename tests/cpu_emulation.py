"""TEST INFRASTRUCTURE — a CPU stand-in for ``backend.CudaKernels`` built on independent torch/NumPy arithmetic,
so the host-side logic of the product (autograd Functions incl. double backward, layer wiring, state_dict
contract, loss graph, gradient bucketing) can be verified without a GPU.  Never used by the product: the only way
to install it is ``backend.set_kernels`` from a test fixture.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import sae_oracle as O


def _nchw(t):
    return t.permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _pad_for(g, x_nchw):
    """explicit zero padding realising (pad_t, pad_l) and whatever bottom/right the output extent needs"""
    need_h = (g.P - 1) * g.stride + g.R
    need_w = (g.Q - 1) * g.stride + g.S
    pb = max(need_h - g.H - g.pad_t, 0)
    pr = max(need_w - g.W - g.pad_l, 0)
    xp = F.pad(x_nchw, (g.pad_l, pr, g.pad_t, pb))
    return xp[:, :, :need_h, :need_w]


def _conv(g, x_nchw, w_krsc):
    return F.conv2d(_pad_for(g, x_nchw), w_krsc.permute(0, 3, 1, 2), stride=g.stride)


class EmulatedKernels:
    name = "cpu-emulation"
    conv_impl = 0
    round_tf32 = False

    def upfirdn2d(self, x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, taps=None):
        if taps is not None:      # the module-supplied factors must reproduce the 2-D kernel
            assert torch.allclose(torch.outer(torch.tensor(taps[0]), torch.tensor(taps[1])).to(kernel.dtype), kernel, atol=1e-6)
        y = O.fir_numpy(_nchw(x).numpy(), kernel.numpy(), (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1))
        return _nhwc(torch.from_numpy(np.ascontiguousarray(y)))

    def bias_act(self, x, bias, ref, act, grad, alpha, scale, noise=None, noise_weight=None):
        t = x
        if bias is not None:
            t = t + bias
        if noise is not None:
            t = t + noise_weight * noise.reshape(*x.shape[:-1], 1)
        if act == 3:
            if grad == 0:
                t = torch.where(t > 0, t, t * alpha)
            elif grad == 1:
                t = torch.where(ref > 0, t, t * alpha)
            else:
                t = torch.zeros_like(t)
        elif grad == 2:
            t = torch.zeros_like(t)
        return t * scale

    def bias_act_backward(self, grad_out, out, alpha, scale, want_bias=True, noise=None, mask=None):
        if mask is not None:
            assert mask.dtype == torch.bool and mask.shape == out.shape and torch.equal(mask, out > 0), \
                "activation mask handed to the backward of a different tensor"
            self.mask_uses = getattr(self, "mask_uses", 0) + 1
        gi = torch.where(out > 0, grad_out, grad_out * alpha) * scale
        c = out.shape[-1]
        gb = gi.reshape(-1, c).sum(0) if want_bias else None
        gnw = None
        if noise is not None:
            gnw = (gi.reshape(-1, c).sum(1) * noise.reshape(-1)).sum().reshape(1)
        return gi, gb, gnw

    def fir_act_backward(self, grad, taps, act_out, pad, alpha, scale, want_bias=True, mask=None):
        kernel = torch.outer(torch.tensor(taps[0]), torch.tensor(taps[1])).to(grad.dtype)
        d = self.upfirdn2d(grad, kernel, 1, 1, 1, 1, *pad)
        gi, gb, _ = self.bias_act_backward(d, act_out, alpha, scale, want_bias=want_bias, mask=mask)
        return gi, gb

    def modulate(self, x, s):
        return x * s[:, None, None, :]

    def modulate_backward(self, dy, x, s):
        return dy * s[:, None, None, :], (dy * x).sum(dim=(1, 2))

    def add_scale(self, a, b, scale):
        return (a + b) * scale if b is not None else a * scale

    def upsample2x_add_scale(self, skip, res, scale):
        up = F.interpolate(_nchw(skip), scale_factor=2, mode="bilinear", align_corners=False)
        return (_nhwc(up) + res) * scale

    def upsample2x_backward(self, dy, scale):
        n, oh, ow, c = dy.shape
        x0 = torch.zeros(n, c, oh // 2, ow // 2, dtype=dy.dtype, requires_grad=True)
        with torch.enable_grad():
            up = F.interpolate(x0, scale_factor=2, mode="bilinear", align_corners=False)
        g, = torch.autograd.grad(up, x0, _nchw(dy).detach())
        return _nhwc(g) * scale

    def pad_channels(self, x, c_out):
        return _nhwc(F.pad(x, (0, 0, 0, 0, 0, c_out - x.shape[1])))

    def reflect_pad(self, x, pads):
        return _nhwc(F.pad(_nchw(x), tuple(pads), mode="reflect"))

    def reflect_pad_backward(self, dy, pads):
        n, oh, ow, c = dy.shape
        pl, pr, pt, pb = pads
        x0 = torch.zeros(n, c, oh - pt - pb, ow - pl - pr, dtype=dy.dtype, requires_grad=True)
        with torch.enable_grad():
            y = F.pad(x0, tuple(pads), mode="reflect")
        g, = torch.autograd.grad(y, x0, _nchw(dy).detach())
        return _nhwc(g)

    def filter_prep(self, w_oihw, scale, want_crsk=True):
        w = w_oihw * scale
        return w.permute(0, 2, 3, 1).contiguous(), (w.permute(1, 2, 3, 0).contiguous() if want_crsk else None)

    def filter_unprep(self, d_krsc, scale):
        return d_krsc.permute(0, 3, 1, 2).contiguous() * scale

    def _epilogue(self, y, bias=None, act=1, alpha=0.2, gain=1.0, noise=None, noise_weight=None, residual=None,
                  res_scale=1.0, round_tf32=None):
        if bias is not None:
            y = y + bias
        if noise is not None:
            y = y + noise_weight * noise.reshape(*y.shape[:-1], 1)
        pos = (y > 0) if act == 3 else None
        if act == 3:
            y = torch.where(y > 0, y, y * alpha)
        y = y * gain
        if residual is not None:
            y = (y + residual) * res_scale
        if pos is not None:
            # stand-in for the activation bit mask of the CUDA kernels (sae_conv_epilogue.act_mask): the callers must hand exactly
            # this object back to the activation backward of exactly this tensor — bias_act_backward checks it
            y._sae_act_mask = pos
        return y

    def conv_fprop(self, x, w_krsc, g, impl=None, prepared=False, **epi):
        y = _nhwc(_conv(g, _nchw(x), w_krsc))
        return self._epilogue(y, **epi)

    def conv_dgrad(self, dy, w_krsc, g, impl=None, w_crsk=None, **epi):
        if w_crsk is not None:
            assert torch.equal(w_crsk, w_krsc.permute(3, 1, 2, 0))
        x0 = torch.zeros(g.N, g.C, g.H, g.W, dtype=dy.dtype, requires_grad=True)
        with torch.enable_grad():
            y = _conv(g, x0, w_krsc.detach())
        dx, = torch.autograd.grad(y, x0, _nchw(dy).detach())
        return self._epilogue(_nhwc(dx), **epi)

    def conv_wgrad(self, dy, x, g, impl=None):
        w0 = torch.zeros(g.K, g.R, g.S, g.C, dtype=dy.dtype, requires_grad=True)
        with torch.enable_grad():
            y = _conv(g, _nchw(x).detach(), w0)
        dw, = torch.autograd.grad(y, w0, _nchw(dy).detach())
        return dw

    def conv_impl_for(self, g, direction):
        return 1

    def bucket_pack(self, *a):
        raise NotImplementedError

    def bucket_unpack(self, *a):
        raise NotImplementedError

    # ------------------------------------------------------------------ crops / Adam
    def crop_gather(self, x, flip, scale, offset, num_crops, size, c_pad, out=None):
        """F.grid_sample formulation of reference util/util.py:323-343, output NHWC zero-padded to ``c_pad`` channels"""
        if out is not None:
            out.copy_(self.crop_gather(x, flip, scale, offset, num_crops, size, c_pad))
            return out
        q = flip.numel()
        lin = torch.linspace(-1.0, 1.0, size, dtype=x.dtype)
        gx = lin.view(1, 1, size, 1).expand(q, size, size, 1)
        gy = lin.view(1, size, 1, 1).expand(q, size, size, 1)
        unit = torch.cat([gx * flip.view(q, 1, 1, 1), gy], dim=3)
        grid = unit * scale.view(q, 1, 1, 2) + offset.view(q, 1, 1, 2)
        xx = x.unsqueeze(1).expand(-1, num_crops, -1, -1, -1).flatten(0, 1)
        crop = F.grid_sample(xx, grid.to(x.dtype), align_corners=False)
        out = torch.zeros(q, size, size, c_pad, dtype=x.dtype)
        out[..., :x.shape[1]] = crop.permute(0, 2, 3, 1)
        return out

    def crop_gather_backward(self, dy, flip, scale, offset, num_crops, c, h, w):
        q = flip.numel()
        x = torch.zeros(q // num_crops, c, h, w, dtype=dy.dtype, requires_grad=True)
        with torch.enable_grad():
            y = self.crop_gather(x, flip, scale, offset, num_crops, dy.shape[2], 4)[..., :c].permute(0, 3, 1, 2)
            gx, = torch.autograd.grad(y, x, dy[:, :c])
        return gx

    def adam_step(self, params, grads, offsets, sizes, exp_avg, exp_avg_sq, steps, lr, beta1, beta2, eps, grad_scale, cache):
        """torch.optim.Adam's documented update, written out per parameter (include/sae_b200.h sae_adam_step)"""
        with torch.no_grad():
            for i, (p, g) in enumerate(zip(params, grads)):
                if g is None:
                    continue
                o, n = int(offsets[i]), int(sizes[i])
                m, v = exp_avg[o:o + n].view_as(p), exp_avg_sq[o:o + n].view_as(p)
                steps[i] += 1
                t = float(steps[i])
                g = g * grad_scale
                m.add_((1 - beta1) * (g - m))
                v.mul_(beta2).add_((1 - beta2) * g * g)
                p.sub_(lr / (1 - beta1 ** t) * m / (v.sqrt() / (1 - beta2 ** t) ** 0.5 + eps))

    def torgb_forward(self, x, s, w, bias, wscale):
        y = torch.einsum("nhwc,nc,oc->nhwo", x, s * wscale, w)
        if bias is not None:
            y = y + bias
        return torch.cat([y, torch.zeros_like(y[..., :1])], dim=3)

    def torgb_backward(self, dy, x, s, w, wscale, want_dx=True, want_gw=True):
        d = dy.permute(0, 2, 3, 1)[..., :3]
        dx = torch.einsum("nhwo,nc,oc->nhwc", d, s * wscale, w) if want_dx else None
        gw = torch.einsum("nhwo,nhwc->noc", d, x) if want_gw else None
        return dx, gw

    def fir_bias_act(self, x, taps, pad, bias, noise, noise_weight, alpha, scale):
        kernel = torch.outer(torch.tensor(taps[0]), torch.tensor(taps[1])).to(x.dtype)
        px0, px1, py0, py1 = pad
        y = self.upfirdn2d(x, kernel, 1, 1, 1, 1, px0, px1, py0, py1)
        out = self.bias_act(y, bias, None, 3, 0, alpha, scale, noise=noise, noise_weight=noise_weight)
        out._sae_act_mask = out > 0
        return out

    # ------------------------------------------------ style-modulated conv, per-sample filters (include/sae_b200.h)
    def conv_modulated_ok(self, g):
        return g.stride == 1 and g.P % 16 == 0 and g.Q % 32 == 0 and g.C % 32 == 0 and g.K % 32 == 0 and g.R <= 3

    def filter_modulate(self, w_krsc, s, want_krsc=True, want_crsk=False):
        wn = w_krsc.unsqueeze(0) * s[:, None, None, None, :]
        return (wn if want_krsc else None), (wn.permute(0, 4, 2, 3, 1).contiguous() if want_crsk else None)

    def conv_fprop_per_sample(self, x, w_nkrsc, g, **epi):
        outs = []
        g1 = type(g)(1, *g.key()[1:])
        for n in range(g.N):
            e = dict(epi)
            if e.get("noise") is not None:
                e["noise"] = e["noise"].reshape(g.N, -1)[n]
            if e.get("residual") is not None:
                e["residual"] = e["residual"][n:n + 1]
            outs.append(self.conv_fprop(x[n:n + 1], w_nkrsc[n], g1, **e))
        y = torch.cat(outs)
        if all(hasattr(o, "_sae_act_mask") for o in outs):
            y._sae_act_mask = torch.cat([o._sae_act_mask for o in outs])
        return y

    def conv_dgrad_per_sample(self, dy, w_ncrsk, g, **epi):
        g1 = type(g)(1, *g.key()[1:])
        return torch.cat([self.conv_dgrad(dy[n:n + 1], w_ncrsk[n].permute(3, 1, 2, 0).contiguous(), g1, **epi) for n in range(g.N)])

    def conv_wgrad_modulated(self, dy, x, s, w_krsc, g):
        g1 = type(g)(1, *g.key()[1:])
        gn = torch.stack([self.conv_wgrad(dy[n:n + 1], x[n:n + 1], g1) for n in range(g.N)])       # [N,K,R,S,C]
        dw = (gn * s[:, None, None, None, :]).sum(0)
        ds = (gn * w_krsc.unsqueeze(0)).sum(dim=(1, 2, 3))
        return dw, ds
