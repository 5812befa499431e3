"""The three helpers of reference ``util/util.py`` that sit on the hot path."""
import torch
import torch.nn.functional as F


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


def apply_random_crop(x, target_size, scale_range, num_crops=1, return_rect=False):
    """Random square crops, random horizontal flip, bilinear resample to ``target_size``
    (reference util/util.py:323-343).  Draw order of the three random tensors (flip, scale, offset) matches the
    reference so a shared RNG seed reproduces the same crops.  Returns [B, num_crops, C, S, S]."""
    b = x.size(0) * num_crops
    flip, scale, offset = draw_crop_parameters(b, scale_range, x.device)
    lin = torch.linspace(-1.0, 1.0, target_size, device=x.device)
    gx = lin.view(1, 1, target_size, 1).expand(b, target_size, target_size, 1)
    gy = lin.view(1, target_size, 1, 1).expand(b, target_size, target_size, 1)
    unit = torch.cat([gx * flip, gy], dim=3)
    x = x.unsqueeze(1).expand(-1, num_crops, -1, -1, -1).flatten(0, 1)
    crop = F.grid_sample(x, (unit * scale + offset).to(x.dtype), align_corners=False)
    return crop.view(b // num_crops, num_crops, crop.size(1), crop.size(2), crop.size(3))


def to_numpy(metric_dict):
    """reference util/util.py:423-429 (one device->host sync per entry)"""
    out = {}
    for k, v in metric_dict.items():
        if "numpy" not in str(type(v)):
            v = v.detach().cpu().mean().numpy()
        out[k] = v
    return out


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
