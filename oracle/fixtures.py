"""ORACLE SUPPORT — deterministic inputs shared by oracle/make_golden.py and tests/ (test infrastructure only)."""
import json
import os

import numpy as np
import torch

from . import sae_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# reduced-capacity option set of the golden network fixtures (same topology as the defaults, 64x64 images)
TINY = dict(num_gpus=0, crop_size=64, batch_size=2, netE_num_downsampling_sp=3, netE_scale_capacity=0.25,
            global_code_ch=128, netG_scale_capacity=0.125, netD_scale_capacity=0.125, netPatchD_scale_capacity=0.5,
            netPatchD_max_nc=32, patch_size=32, patch_num_crops=2)


def rnd(seed, *shape, dtype=torch.float64):
    """standard-normal tensor from NumPy's legacy generator (bit-stable across machines)"""
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape)).to(dtype)


def perturbed_state_dict(opt, dtype=torch.float64, param_seed=7, bias_seed=11):
    """init_state_dict + non-zero biases and noise weights so that parity checks exercise every term"""
    sd = O.init_state_dict(opt, seed=param_seed, dtype=dtype)
    rs = np.random.RandomState(bias_seed)
    for k in sd:
        if k.endswith(".bias") and "modulation" not in k:
            sd[k] = sd[k] + torch.from_numpy(rs.standard_normal(tuple(sd[k].shape)) * 0.1).to(dtype)
        if k.endswith("noise.weight"):
            sd[k] = sd[k] + 0.05
    return sd


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, arrays


def rel_err(a, b):
    """max|a-b| / max|b|  — the tolerance metric of BASELINE.json ("1e-3 rel"), SURVEY.md §9.5"""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
