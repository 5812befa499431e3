"""TEST INFRASTRUCTURE — "shadow" kernel set: every call goes to the CUDA kernels AND, on CPU copies of the same
inputs in fp64, to the oracle-backed emulation; the per-call error is recorded.  Running a real network step under
it checks every kernel configuration the networks actually produce (shapes, pads, strides, epilogues) and points
at the exact call that deviates."""
import torch

from swapping_autoencoder_pytorch_b200 import backend
from tests.cpu_emulation import EmulatedKernels


def _cpu(v):
    if torch.is_tensor(v):
        return v.detach().double().cpu()
    return v


def _err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class ShadowKernels:
    name = "shadow"

    def __init__(self):
        self.real = backend.CudaKernels()
        self.emu = EmulatedKernels()
        self.log = []          # (method, description, error)

    @property
    def conv_impl(self):
        return self.real.conv_impl

    @conv_impl.setter
    def conv_impl(self, v):
        self.real.conv_impl = v

    @property
    def round_tf32(self):
        return self.real.round_tf32

    def _both(self, name, desc, args, kwargs):
        out = getattr(self.real, name)(*args, **kwargs)
        kw = {k: _cpu(v) for k, v in kwargs.items() if k != "round_tf32"}
        if "taps" in kw:
            kw["taps"] = kwargs["taps"]
        ref = getattr(self.emu, name)(*[_cpu(a) for a in args], **kw)
        outs = out if isinstance(out, tuple) else (out,)
        refs = ref if isinstance(ref, tuple) else (ref,)
        for i, (o, r) in enumerate(zip(outs, refs)):
            if o is None:
                continue
            self.log.append((name, "%s out%d" % (desc, i), _err(o, r)))
        return out

    def upfirdn2d(self, x, kernel, *cfg, taps=None):
        return self._both("upfirdn2d", "x%s k%s cfg%s sep%s" % (tuple(x.shape), tuple(kernel.shape), cfg, taps is not None),
                          (x, kernel) + cfg, dict(taps=taps))

    def bias_act(self, x, bias, ref, act, grad, alpha, scale, noise=None, noise_weight=None):
        return self._both("bias_act", "x%s act%d grad%d noise%s" % (tuple(x.shape), act, grad, noise is not None),
                          (x, bias, ref, act, grad, alpha, scale), dict(noise=noise, noise_weight=noise_weight))

    def bias_act_backward(self, grad_out, out, alpha, scale, want_bias=True, noise=None, mask=None):
        # (the activation bit mask goes to the real kernel only; the emulation reads ``out``: a wrong mask shows up as an error)
        res = self.real.bias_act_backward(grad_out, out, alpha, scale, want_bias=want_bias, noise=noise, mask=mask)
        ref = self.emu.bias_act_backward(_cpu(grad_out), _cpu(out), alpha, scale, want_bias=want_bias, noise=_cpu(noise))
        for i, (o, r) in enumerate(zip(res, ref)):
            if o is not None:
                self.log.append(("bias_act_backward", "x%s noise%s mask%s out%d" % (tuple(out.shape), noise is not None, mask is not None, i),
                                 _err(o, r)))
        return res

    def fir_act_backward(self, grad, taps, act_out, pad, alpha, scale, want_bias=True, mask=None):
        out = self.real.fir_act_backward(grad, taps, act_out, pad, alpha, scale, want_bias=want_bias, mask=mask)
        if out is None:           # shape outside the fused kernel: the caller issues the two separate (shadowed) calls
            return None
        ref = self.emu.fir_act_backward(_cpu(grad), taps, _cpu(act_out), pad, alpha, scale, want_bias=want_bias)
        for i, (o, r) in enumerate(zip(out, ref)):
            if o is not None:
                self.log.append(("fir_act_backward", "g%s pad%s out%d" % (tuple(grad.shape), pad, i), _err(o, r)))
        return out

    def fir_bias_act(self, x, taps, pad, bias, noise, noise_weight, alpha, scale):
        out = self.real.fir_bias_act(x, taps, pad, bias, noise, noise_weight, alpha, scale)
        if out is None:           # shape outside the fused kernel: the caller issues the two separate (shadowed) calls
            return None
        ref = self.emu.fir_bias_act(_cpu(x), taps, pad, _cpu(bias), _cpu(noise), _cpu(noise_weight), alpha, scale)
        self.log.append(("fir_bias_act", "x%s pad%s noise%s" % (tuple(x.shape), pad, noise is not None), _err(out, ref)))
        return out

    def torgb_forward(self, x, s, w, bias, wscale):
        return self._both("torgb_forward", "x%s" % (tuple(x.shape),), (x, s, w, bias, wscale), {})

    def torgb_backward(self, dy, x, s, w, wscale, want_dx=True, want_gw=True):
        return self._both("torgb_backward", "x%s" % (tuple(x.shape),), (dy, x, s, w, wscale), dict(want_dx=want_dx, want_gw=want_gw))

    def crop_gather(self, x, flip, scale, offset, num_crops, size, c_pad, out=None):
        res = self.real.crop_gather(x, flip, scale, offset, num_crops, size, c_pad, out=out)
        ref = self.emu.crop_gather(_cpu(x), _cpu(flip), _cpu(scale), _cpu(offset), num_crops, size, c_pad)
        self.log.append(("crop_gather", "x%s -> %d" % (tuple(x.shape), size), _err(res, ref)))
        return res

    def crop_gather_backward(self, dy, flip, scale, offset, num_crops, c, h, w):
        return self._both("crop_gather_backward", "dy%s" % (tuple(dy.shape),), (dy, flip, scale, offset, num_crops, c, h, w), {})

    def conv_modulated_ok(self, g):
        return self.real.conv_modulated_ok(g)

    def filter_modulate(self, w_krsc, s, want_krsc=True, want_crsk=False):
        return self._both("filter_modulate", "w%s" % (tuple(w_krsc.shape),), (w_krsc, s), dict(want_krsc=want_krsc, want_crsk=want_crsk))

    def conv_fprop_per_sample(self, x, w, g, **epi):
        return self._both("conv_fprop_per_sample", "g%s epi%s" % (g.key(), sorted(epi)), (x, w, g), dict(epi))

    def conv_dgrad_per_sample(self, dy, w, g, **epi):
        return self._both("conv_dgrad_per_sample", "g%s" % (g.key(),), (dy, w, g), dict(epi))

    def conv_wgrad_modulated(self, dy, x, s, w, g):
        return self._both("conv_wgrad_modulated", "g%s" % (g.key(),), (dy, x, s, w, g), {})

    def adam_step(self, *a, **kw):
        return self.real.adam_step(*a, **kw)         # checked against torch.optim.Adam in tests/test_gpu_train_ops.py

    def modulate(self, x, s):
        return self._both("modulate", "x%s" % (tuple(x.shape),), (x, s), {})

    def modulate_backward(self, dy, x, s):
        return self._both("modulate_backward", "x%s" % (tuple(x.shape),), (dy, x, s), {})

    def add_scale(self, a, b, scale):
        return self._both("add_scale", "x%s" % (tuple(a.shape),), (a, b, scale), {})

    def upsample2x_add_scale(self, skip, res, scale):
        return self._both("upsample2x_add_scale", "x%s" % (tuple(skip.shape),), (skip, res, scale), {})

    def upsample2x_backward(self, dy, scale):
        return self._both("upsample2x_backward", "x%s" % (tuple(dy.shape),), (dy, scale), {})

    def pad_channels(self, x, c_out):
        return self._both("pad_channels", "x%s -> %d" % (tuple(x.shape), c_out), (x, c_out), {})

    def reflect_pad(self, x, pads):
        return self._both("reflect_pad", "x%s pads%s" % (tuple(x.shape), pads), (x, pads), {})

    def reflect_pad_backward(self, dy, pads):
        return self._both("reflect_pad_backward", "x%s pads%s" % (tuple(dy.shape), pads), (dy, pads), {})

    def _impl_name(self, g, d, impl):
        return impl if impl is not None else self.real.conv_impl_for(g, d)

    def filter_prep(self, w, scale, want_crsk=True):
        return self._both("filter_prep", "w%s" % (tuple(w.shape),), (w, scale), dict(want_crsk=want_crsk))

    def filter_unprep(self, d, scale):
        return self._both("filter_unprep", "w%s" % (tuple(d.shape),), (d, scale), {})

    def conv_fprop(self, x, w, g, impl=None, prepared=False, **epi):
        return self._both("conv_fprop", "g%s impl%s epi%s" % (g.key(), self._impl_name(g, 0, impl), sorted(epi)), (x, w, g),
                          dict(epi, impl=impl, prepared=prepared))

    def conv_dgrad(self, dy, w, g, impl=None, w_crsk=None, **epi):
        return self._both("conv_dgrad", "g%s impl%s" % (g.key(), self._impl_name(g, 1, impl)), (dy, w, g),
                          dict(epi, impl=impl, w_crsk=w_crsk))

    def conv_wgrad(self, dy, x, g, impl=None):
        return self._both("conv_wgrad", "g%s impl%s" % (g.key(), self._impl_name(g, 2, impl)), (dy, x, g), dict(impl=impl))

    def conv_impl_for(self, g, d):
        return self.real.conv_impl_for(g, d)

    def bucket_pack(self, *a):
        return self.real.bucket_pack(*a)

    def bucket_unpack(self, *a):
        return self.real.bucket_unpack(*a)

    def worst(self, n=12):
        return sorted(self.log, key=lambda t: -t[2])[:n]

    def failures(self, tol_conv=1.5e-3, tol_other=5e-4):
        bad = []
        for name, desc, e in self.log:
            tol = tol_conv if name.startswith("conv") else tol_other
            if not (e < tol):
                bad.append((name, desc, e))
        return bad
