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

    def bias_act_backward(self, grad_out, out, alpha, scale, want_bias=True, noise=None):
        gi = torch.where(out > 0, grad_out, grad_out * alpha) * scale
        c = out.shape[-1]
        gb = gi.reshape(-1, c).sum(0) if want_bias else None
        gnw = None
        if noise is not None:
            gnw = (gi.reshape(-1, c).sum(1) * noise.reshape(-1)).sum().reshape(1)
        return gi, gb, gnw

    def fir_act_backward(self, grad, taps, act_out, pad, alpha, scale, want_bias=True):
        kernel = torch.outer(torch.tensor(taps[0]), torch.tensor(taps[1])).to(grad.dtype)
        d = self.upfirdn2d(grad, kernel, 1, 1, 1, 1, *pad)
        gi, gb, _ = self.bias_act_backward(d, act_out, alpha, scale, want_bias=want_bias)
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
        if act == 3:
            y = torch.where(y > 0, y, y * alpha)
        y = y * gain
        if residual is not None:
            y = (y + residual) * res_scale
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
