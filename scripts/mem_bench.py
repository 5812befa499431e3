#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels at hot-path shapes (CUDA-event timed, L2 flushed).
GB/s are ALGORITHMIC bytes: FIR 4*(N_in+N_out), bias-act fwd 8N / bwd 12N, modulate 8N, add_scale 12N."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swapping_autoencoder_pytorch_b200 import backend  # noqa: E402


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
    k4 = torch.tensor([1., 3., 3., 1.], device=dev)
    k4 = torch.outer(k4, k4) / 64
    k3 = torch.tensor([1., 2., 1.], device=dev)
    k3 = torch.outer(k3, k3) / 16
    print("%-46s %9s %9s" % ("kernel / shape", "ms", "GB/s"))
    for n, h, c in ((32, 256, 128), (32, 128, 256), (32, 64, 512), (256, 128, 32), (32, 256, 32)):
        x = torch.randn(n, h, h, c, device=dev)
        t4, t3 = (0.125, 0.375, 0.375, 0.125), (0.25, 0.5, 0.25)
        for name, kern, pad, tp in (("fir4 pad(2,2)", k4, (2, 2), t4), ("fir4 pad(1,1)", k4, (1, 1), t4), ("fir3 pad(0,0)", k3, (0, 0), t3)):
            ms = timeit(lambda: k.upfirdn2d(x, kern, 1, 1, 1, 1, pad[0], pad[1], pad[0], pad[1], taps=(tp, tp)), flush)
            oh = h + pad[0] + pad[1] - kern.shape[0] + 1
            gb = 4.0 * (x.numel() + n * oh * oh * c) / 1e9
            print("%-46s %9.3f %9.0f" % ("%s [%d,%d,%d,%d]" % (name, n, h, h, c), ms, gb / ms * 1e3))
        # decimating blur of the skip branches (down 2, pad (1,1)) and its adjoint (zero-insert x2)
        ms = timeit(lambda: k.upfirdn2d(x, k4, 1, 1, 2, 2, 1, 1, 1, 1, taps=(t4, t4)), flush)
        oh = (h + 2 - 4) // 2 + 1
        print("%-46s %9.3f %9.0f" % ("fir4 down2 [%d,%d,%d,%d]" % (n, h, h, c), ms, 4.0 * (x.numel() + n * oh * oh * c) / 1e9 / ms * 1e3))
        xs = torch.randn(n, oh, oh, c, device=dev)
        ms = timeit(lambda: k.upfirdn2d(xs, k4, 2, 2, 1, 1, 2, h - oh * 2 + 1, 2, h - oh * 2 + 1, taps=(t4, t4)), flush)
        print("%-46s %9.3f %9.0f" % ("fir4 up2 (adjoint) -> [%d,%d,%d,%d]" % (n, h, h, c), ms, 4.0 * (x.numel() + xs.numel()) / 1e9 / ms * 1e3))
        # blur adjoint fused with the activation backward
        act = torch.randn(n, h, h, c, device=dev)
        gsrc = torch.randn(n, h + 1, h + 1, c, device=dev)
        ms = timeit(lambda: k.fir_act_backward(gsrc, (t4, t4), act, (1, 1, 1, 1), 0.2, 1.414), flush)
        print("%-46s %9.3f %9.0f" % ("fir4 + act bwd fused -> [%d,%d,%d,%d]" % (n, h, h, c), ms, 4.0 * (gsrc.numel() + 2 * act.numel()) / 1e9 / ms * 1e3))
        del xs, act, gsrc
        b = torch.randn(c, device=dev)
        y = torch.randn_like(x)
        ms = timeit(lambda: k.bias_act(x, b, None, 3, 0, 0.2, 1.414), flush)
        print("%-46s %9.3f %9.0f" % ("bias_act fwd [%d,%d,%d,%d]" % (n, h, h, c), ms, 8.0 * x.numel() / 1e9 / ms * 1e3))
        ms = timeit(lambda: k.bias_act_backward(x, y, 0.2, 1.414), flush)
        print("%-46s %9.3f %9.0f" % ("bias_act bwd [%d,%d,%d,%d]" % (n, h, h, c), ms, 12.0 * x.numel() / 1e9 / ms * 1e3))
        s = torch.randn(n, c, device=dev)
        ms = timeit(lambda: k.modulate(x, s), flush)
        print("%-46s %9.3f %9.0f" % ("modulate [%d,%d,%d,%d]" % (n, h, h, c), ms, 8.0 * x.numel() / 1e9 / ms * 1e3))
        ms = timeit(lambda: k.modulate_backward(x, y, s), flush)
        print("%-46s %9.3f %9.0f" % ("modulate bwd [%d,%d,%d,%d]" % (n, h, h, c), ms, 12.0 * x.numel() / 1e9 / ms * 1e3))
        ms = timeit(lambda: k.add_scale(x, y, 0.7), flush)
        print("%-46s %9.3f %9.0f" % ("add_scale [%d,%d,%d,%d]" % (n, h, h, c), ms, 12.0 * x.numel() / 1e9 / ms * 1e3))
        if h <= 128:
            r = torch.randn(n, 2 * h, 2 * h, c, device=dev)
            ms = timeit(lambda: k.upsample2x_add_scale(x, r, 0.7), flush)
            print("%-46s %9.3f %9.0f" % ("upsample2x_add [%d,%d,%d,%d]" % (n, h, h, c), ms, (4.0 * x.numel() + 8.0 * r.numel()) / 1e9 / ms * 1e3))
            ms = timeit(lambda: k.upsample2x_backward(r, 0.7), flush)
            print("%-46s %9.3f %9.0f" % ("upsample2x_bwd [%d,%d,%d,%d]" % (n, h, h, c), ms, (4.0 * x.numel() + 4.0 * r.numel()) / 1e9 / ms * 1e3))
            del r
        del x, y


if __name__ == "__main__":
    main()
