#!/usr/bin/env python
"""Per-shape throughput of the conv kernels (CUDA-event timed, L2 flushed between iterations).
usage: python scripts/conv_bench.py [--impl 0|1|2] [--dirs fprop,dgrad,wgrad] [--only SUBSTR] [--iters N] [--batch B]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swapping_autoencoder_pytorch_b200 import backend  # noqa: E402
from swapping_autoencoder_pytorch_b200.backend import make_geom  # noqa: E402

# (name, H, C, K, R, stride, pad, images-per-32-batch multiplier)  — SURVEY.md Appendix A, the layers carrying the FLOPs
SHAPES = [
    ("D/G 128->128 @256 s1", 256, 128, 128, 3, 1, 1, 1.0),
    ("D/G 256->256 @128 s1", 128, 256, 256, 3, 1, 1, 1.0),
    ("D/G 512->512 @64 s1", 64, 512, 512, 3, 1, 1, 1.0),
    ("D 512->512 @32 s1", 32, 512, 512, 3, 1, 1, 1.0),
    ("D 128->256 @257 s2", 257, 128, 256, 3, 2, 0, 1.0),
    ("D 256->512 @129 s2", 129, 256, 512, 3, 2, 0, 1.0),
    ("D 512->512 @65 s2", 65, 512, 512, 3, 2, 0, 1.0),
    ("(no strips) 128->256 @256 s2", 256, 128, 256, 3, 2, 0, 1.0),
    ("D skip 128->256 @255 1x1 s2", 255, 128, 256, 1, 2, 0, 1.0),
    ("G skip 512->256 @64 1x1", 64, 512, 256, 1, 1, 0, 1.0),
    ("G convT 256->128 @128 (as dgrad of s2)", 257, 128, 256, 3, 2, 0, 1.0),
    ("FromRGB 32->128 @256 1x1", 256, 32, 128, 1, 1, 0, 1.0),
    ("D skip 128->256 @128 1x1 (decimated)", 128, 128, 256, 1, 1, 0, 1.0),
    ("D skip 256->512 @64 1x1 (decimated)", 64, 256, 512, 1, 1, 0, 1.0),
    ("Dpatch skip 32->64 @64 1x1 (decimated)", 64, 32, 64, 1, 1, 0, 8.0),
    ("Dpatch 32->32 @128 s1", 128, 32, 32, 3, 1, 1, 8.0),
    ("Dpatch 64->64 @64 s1", 64, 64, 64, 3, 1, 1, 8.0),
    ("E 32->32 @258 s1 p0", 258, 32, 32, 3, 1, 0, 1.0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--dirs", default="fprop,dgrad,wgrad")
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    k = backend.kernels()
    k.conv_impl = args.impl
    dev = torch.device("cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    print("%-42s %-6s %5s %9s %9s %8s" % ("shape", "dir", "impl", "ms", "TFLOP/s", "GB/s"))
    for name, h, c, kk, r, stride, pad, mult in SHAPES:
        if args.only and args.only not in name:
            continue
        n = max(int(args.batch * mult), 1)
        g = make_geom(n, h, h, c, kk, r, r, stride, pad, pad)
        x = torch.randn(n, h, h, c, device=dev)
        w = torch.randn(kk, r, r, c, device=dev) / (c * r * r) ** 0.5
        dy = torch.randn(n, g.P, g.Q, kk, device=dev)
        flops = 2.0 * n * g.P * g.Q * kk * r * r * c
        nbytes = 4.0 * (x.numel() + dy.numel() + w.numel())
        for d in args.dirs.split(","):
            fn = {"fprop": lambda: k.conv_fprop(x, w, g), "dgrad": lambda: k.conv_dgrad(dy, w, g),
                  "wgrad": lambda: k.conv_wgrad(dy, x, g)}[d]
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            impl = k.conv_impl_for(g, {"fprop": 0, "dgrad": 1, "wgrad": 2}[d]) if args.impl == 0 else args.impl
            print("%-42s %-6s %5d %9.3f %9.1f %8.0f" % (name, d, impl, ms, flops / ms / 1e9, nbytes / ms / 1e6), flush=True)
        del x, w, dy


if __name__ == "__main__":
    main()
