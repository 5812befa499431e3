"""The three helpers of reference ``util/util.py`` that sit on the hot path."""
import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import backend


def normalize(v):
    """L2-normalise over dim 1 (reference util/util.py:18-22)."""
    if isinstance(v, list):
        return [normalize(t) for t in v]
    return v * torch.rsqrt(v.square().sum(dim=1, keepdim=True) + 1e-8)


def draw_crop_parameters(b, scale_range, device):
    """The three random tensors of a crop batch, drawn in the reference's order (flip, scale, offset)."""
    lo, hi = scale_range
    flip = torch.round(torch.rand(b, 1, 1, 1, device=device)) * 2 - 1.0
    scale = torch.rand(b, 1, 1, 2, device=device) * (hi - lo) + lo
    offset = (torch.rand(b, 1, 1, 2, device=device) * 2 - 1) * (1 - scale)
    return flip, scale, offset


class _CropGather(Function):
    """The crop resampler as ONE kernel (reference util/util.py:323-343 builds an affine grid and calls F.grid_sample): writes
    the crops as a channels-last tensor zero-padded to 32 channels — the layout the first patch-discriminator convolution
    reads, so no separate channel-pad / layout pass follows (stylegan2_op/conv.py::_pad4 recognises the buffer) — and returns
    the logical [Q, C, S, S] view of it.  Linear in x; backward is the adjoint kernel in gather form (no atomics).  The crops
    of R1 are detached leaves, the generator step differentiates once: once-differentiable."""

    PAD = 32

    @staticmethod
    def forward(ctx, x, flip, scale, offset, num_crops, size):
        k = backend.kernels()
        flip, scale, offset = flip.reshape(-1).contiguous(), scale.reshape(-1, 2).contiguous(), offset.reshape(-1, 2).contiguous()
        buf = k.crop_gather(x, flip, scale, offset, num_crops, size, _CropGather.PAD)        # [Q, S, S, 32]
        buf._sae_zero_padded = _CropGather.PAD       # channels C.. are zeros: see conv._pad4
        ctx.save_for_backward(flip, scale, offset)
        ctx.meta = (num_crops, tuple(x.shape))
        return buf.permute(0, 3, 1, 2)[:, :x.shape[1]]

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        flip, scale, offset = ctx.saved_tensors
        num_crops, (b, c, h, w) = ctx.meta
        return backend.kernels().crop_gather_backward(dy, flip, scale, offset, num_crops, c, h, w), None, None, None, None, None


class _CropGatherMulti(Function):
    """_CropGather over several source batches into ONE output buffer (one launch per source, no concatenation pass): the
    batched discriminator passes feed the patch discriminator the crops of (real, real, mix) as a single batch."""

    @staticmethod
    def forward(ctx, num_crops, size, *args):
        k = backend.kernels()
        srcs = [args[i:i + 4] for i in range(0, len(args), 4)]
        qs = [s[1].numel() for s in srcs]
        c = srcs[0][0].shape[1]
        buf = torch.empty((sum(qs), size, size, _CropGather.PAD), device=srcs[0][0].device, dtype=srcs[0][0].dtype)
        saved, o = [], 0
        for (x, flip, scale, offset), q in zip(srcs, qs):
            flip, scale, offset = flip.reshape(-1).contiguous(), scale.reshape(-1, 2).contiguous(), offset.reshape(-1, 2).contiguous()
            k.crop_gather(x, flip, scale, offset, num_crops, size, _CropGather.PAD, out=buf[o:o + q])
            saved += [flip, scale, offset]
            o += q
        buf._sae_zero_padded = _CropGather.PAD
        ctx.save_for_backward(*saved)
        ctx.meta = (num_crops, [tuple(s[0].shape) for s in srcs], qs)
        return buf.permute(0, 3, 1, 2)[:, :c]

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        num_crops, shapes, qs = ctx.meta
        saved = ctx.saved_tensors
        grads, o = [None, None], 0
        for i, ((b, c, h, w), q) in enumerate(zip(shapes, qs)):
            dx = None
            if ctx.needs_input_grad[2 + 4 * i]:
                flip, scale, offset = saved[3 * i:3 * i + 3]
                dx = backend.kernels().crop_gather_backward(dy[o:o + q], flip, scale, offset, num_crops, c, h, w)
            grads += [dx, None, None, None]
            o += q
        return tuple(grads)


def apply_random_crops_multi(xs, target_size, scale_range, num_crops=1):
    """``[apply_random_crop(x, ...) for x in xs]`` concatenated along the batch — same random draws in the same order — written
    by the crop kernel straight into one zero-padded channels-last buffer.  Returns ([sum B, num_crops, C, S, S], [B_i])."""
    args = []
    for x in xs:
        args += [x, *draw_crop_parameters(x.size(0) * num_crops, scale_range, x.device)]
    crop = _CropGatherMulti.apply(num_crops, target_size, *args)
    return crop.view(crop.size(0) // num_crops, num_crops, crop.size(1), crop.size(2), crop.size(3)), [x.size(0) for x in xs]


def apply_random_crop(x, target_size, scale_range, num_crops=1, return_rect=False):
    """Random square crops, random horizontal flip, bilinear resample to ``target_size``
    (reference util/util.py:323-343).  Draw order of the three random tensors (flip, scale, offset) matches the
    reference so a shared RNG seed reproduces the same crops.  Returns [B, num_crops, C, S, S]."""
    b = x.size(0) * num_crops
    flip, scale, offset = draw_crop_parameters(b, scale_range, x.device)
    if x.size(1) <= 4 and target_size >= 2 and hasattr(backend.kernels(), "crop_gather"):
        crop = _CropGather.apply(x, flip, scale, offset, num_crops, target_size)
        return crop.view(b // num_crops, num_crops, crop.size(1), crop.size(2), crop.size(3))
    lin = torch.linspace(-1.0, 1.0, target_size, device=x.device)
    gx = lin.view(1, 1, target_size, 1).expand(b, target_size, target_size, 1)
    gy = lin.view(1, target_size, 1, 1).expand(b, target_size, target_size, 1)
    unit = torch.cat([gx * flip, gy], dim=3)
    x = x.unsqueeze(1).expand(-1, num_crops, -1, -1, -1).flatten(0, 1)
    crop = F.grid_sample(x, (unit * scale + offset).to(x.dtype), align_corners=False)
    return crop.view(b // num_crops, num_crops, crop.size(1), crop.size(2), crop.size(3))


class LazyLosses(dict):
    """The dict ``to_numpy`` returns, with the device -> host wait deferred to the first time a VALUE is looked at.  The
    reference's ``to_numpy`` (util/util.py:423-429) blocks the host on every entry of every half-step; here all means are
    gathered by one small device op, copied to pinned memory asynchronously, and the host runs ahead to issue the next
    half-step.  Keys (``"D_R1" in losses``, ``len``, iteration) never wait."""

    def __init__(self, keys, host, event, eager=None):
        super().__init__()
        self._pending = (list(keys), host, event)
        for k in keys:
            dict.__setitem__(self, k, None)
        if eager:
            for k, v in eager.items():
                dict.__setitem__(self, k, v)

    def _resolve(self):
        pending, self._pending = self._pending, None
        if pending is not None:
            keys, host, event = pending
            event.synchronize()
            arr = host.numpy()
            for i, k in enumerate(keys):
                dict.__setitem__(self, k, arr[i].copy())

    def __getitem__(self, k):
        self._resolve()
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        self._resolve()
        return dict.get(self, k, default)

    def items(self):
        self._resolve()
        return dict.items(self)

    def values(self):
        self._resolve()
        return dict.values(self)

    def copy(self):
        self._resolve()
        return dict(self)

    def __repr__(self):
        self._resolve()
        return dict.__repr__(self)


def to_numpy(metric_dict, lazy=False):
    """reference util/util.py:423-429: every entry reduced to its mean as a NumPy scalar.  ``lazy`` (extension, CUDA only): one
    gather + one asynchronous copy for the whole dict, the wait deferred until a value is read (``LazyLosses``)."""
    dev = [v for v in metric_dict.values() if torch.is_tensor(v) and v.is_cuda]
    if lazy and dev and len(dev) == sum(1 for v in metric_dict.values() if torch.is_tensor(v)):
        keys = [k for k, v in metric_dict.items() if torch.is_tensor(v)]
        with torch.no_grad():
            flat = torch.stack([metric_dict[k].detach().float().mean() for k in keys])
            host = torch.empty(len(keys), dtype=torch.float32, pin_memory=True)
            host.copy_(flat, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        return LazyLosses(keys, host, event, {k: v for k, v in metric_dict.items() if not torch.is_tensor(v)})
    out = {}
    for k, v in metric_dict.items():
        if "numpy" not in str(type(v)):
            v = v.detach().cpu().mean().numpy()
        out[k] = v
    return out


def visualize_spatial_code(sp):
    """structure code -> 3-channel picture for the snapshot's "layout" entry (reference util/util.py:231-255: 3-component
    PCA over all code vectors of the batch, rescaled to [-1, 1]).  Visualisation only; computed with an SVD in torch, signs
    fixed like scikit-learn's svd_flip so the picture does not flip between calls."""
    if sp.size(1) <= 2:
        sp = sp.repeat([1, 3, 1, 1])[:, :3]
    if sp.size(1) == 3:
        return sp
    b, c, h, w = sp.shape
    x = sp.detach().permute(0, 2, 3, 1).reshape(-1, c).double()
    x = x - x.mean(dim=0, keepdim=True)
    try:
        u, s_, _ = torch.linalg.svd(x, full_matrices=False)
    except RuntimeError:
        return torch.zeros(b, 3, h, w, device=sp.device, dtype=sp.dtype)
    u = u[:, :3] * s_[:3]
    sign = torch.sign(u.gather(0, u.abs().argmax(dim=0, keepdim=True)))
    z = (u * sign).reshape(b, h, w, 3).permute(0, 3, 1, 2)
    z = (z - z.min()) / (z.max() - z.min()).clamp_min(1e-30) * 2 - 1
    return z.to(sp.dtype)


def resize2d_tensor(x, size_or_tensor_of_size):
    """reference util/util.py:457-469"""
    size = size_or_tensor_of_size.size() if torch.is_tensor(size_or_tensor_of_size) else size_or_tensor_of_size
    return F.interpolate(x, tuple(size)[-2:], mode='bilinear', align_corners=False)


def gan_loss(pred, should_be_classified_as_real):
    """Non-saturating logistic loss per sample (reference models/networks/loss.py:11-16)."""
    bs = pred.size(0)
    z = -pred if should_be_classified_as_real else pred
    return F.softplus(z).view(bs, -1).mean(dim=1)


def str2bool(v):
    """argparse helper (reference util/util.py str2bool)"""
    if isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise ValueError('Boolean value expected.')
