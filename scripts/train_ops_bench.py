#!/usr/bin/env python
"""Achieved HBM bandwidth of this round's bandwidth-bound kernels at hot-path shapes (CUDA-event timed, L2 flushed between
iterations).  GB/s are ALGORITHMIC bytes: ToRGB fwd 4 N_x (+ 16 B / pixel out), bwd 8 N_x; crop forward 128 B written per crop
pixel; Adam 16 B read + 12 B written per element; FIR + bias + activation 4 (N_in + N_out)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swapping_autoencoder_pytorch_b200 import backend  # noqa: E402
from swapping_autoencoder_pytorch_b200.optimizer import MultiTensorAdam  # noqa: E402


def timeit(fn, flush, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    k = backend.kernels()
    dev = torch.device("cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    print("%-58s %9s %9s" % ("kernel / shape", "ms", "GB/s"))

    def row(name, ms, nbytes):
        print("%-58s %9.3f %9.0f" % (name, ms, nbytes / ms / 1e6), flush=True)

    # ToRGB 128 -> 3 at 256^2, batch 32 and 16
    for n in (32, 16):
        x = torch.randn(n, 256, 256, 128, device=dev)
        s, w, b = torch.randn(n, 128, device=dev), torch.randn(3, 128, device=dev), torch.randn(3, device=dev)
        dy = torch.randn(n, 3, 256, 256, device=dev)
        row("torgb_forward  %dx256x256x128" % n, timeit(lambda: k.torgb_forward(x, s, w, b, 0.088), flush), 4.0 * x.numel() + 16.0 * n * 65536)
        row("torgb_backward %dx256x256x128 (dx + G)" % n, timeit(lambda: k.torgb_backward(dy, x, s, w, 0.088), flush), 8.0 * x.numel())
        del x, dy
    # crops: 32 images x 8 crops of 128^2 from 256^2
    img = torch.randn(32, 3, 256, 256, device=dev)
    q = 256
    flip = torch.ones(q, device=dev)
    scale = torch.rand(q, 2, device=dev) * 0.125 + 0.125
    offset = (torch.rand(q, 2, device=dev) * 2 - 1) * (1 - scale)
    row("crop_gather 256 crops 128^2 -> [256,128,128,32]", timeit(lambda: k.crop_gather(img, flip, scale, offset, 8, 128, 32), flush),
        128.0 * q * 128 * 128)
    dyc = torch.randn(q, 128, 128, 32, device=dev).permute(0, 3, 1, 2)[:, :3]
    row("crop_gather_backward (gather form) -> [32,3,256,256]", timeit(lambda: k.crop_gather_backward(dyc, flip, scale, offset, 8, 3, 256, 256), flush),
        12.0 * q * 128 * 128 + 4.0 * img.numel())
    # Adam over a 55 M-parameter group in 224 tensors
    sizes = [512 * 512 * 9] * 18 + [256 * 256 * 9] * 8 + [2048 * 2048] + [512] * 197
    params = [torch.randn(sz, device=dev).requires_grad_() for sz in sizes]
    for p in params:
        p.grad = torch.randn_like(p)
    opt = MultiTensorAdam(params, lr=0.002, betas=(0.0, 0.99))
    tot = sum(sizes)
    row("adam_step %d tensors, %.1f M elements" % (len(sizes), tot / 1e6), timeit(opt.step, flush), 28.0 * tot)
    del params, opt
    # FIR + noise + bias + activation (the blur behind the generator's transposed conv), 257^2 -> 256^2 x 128 ch, batch 32
    for n, h, c in ((32, 257, 128), (32, 129, 256)):
        x = torch.randn(n, h, h, c, device=dev)
        t = (0.25, 0.75, 0.75, 0.25)
        bias, noise, nw = torch.randn(c, device=dev), torch.randn(n * (h - 1) * (h - 1), device=dev), torch.ones(1, device=dev)
        row("fir_bias_act %dx%dx%dx%d" % (n, h, h, c), timeit(lambda: k.fir_bias_act(x, (t, t), (1, 1, 1, 1), bias, noise, nw, 0.2, 1.4), flush),
            4.0 * (x.numel() + n * (h - 1) * (h - 1) * c))
        del x
    # per-sample filters of the modulated convolution
    w = torch.randn(128, 3, 3, 128, device=dev)
    s = torch.randn(32, 128, device=dev)
    row("filter_modulate 32 x [128,3,3,128] (both layouts)", timeit(lambda: k.filter_modulate(w, s, True, True), flush), 8.0 * 32 * w.numel())


if __name__ == "__main__":
    main()
