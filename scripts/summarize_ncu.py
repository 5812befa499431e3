#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown).
usage: python scripts/summarize_ncu.py gpurun_out/launches.csv > profiles/launches_summary.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0.0, 0])
    for row in rows:
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row.get("Metric Unit", "ns")
        v *= {"us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
        name = row["Kernel Name"]
        if "sae::" in name:
            short = re.sub(r"\(.*", "", name)
            short = re.sub(r"^void ", "", short)
        else:
            m = re.search(r"(\w+Functor\w*|\w+_kernel_cuda|upsample\w+|grid_sampler\w+|reflection\w+|CatArray\w*|reduce_kernel|"
                          r"elementwise_kernel|vectorized_elementwise_kernel|index_elementwise_kernel|multi_tensor\w+)", name)
            short = "torch: " + (m.group(1) if m else name[:50])
        tot[short][0] += v
        tot[short][1] += 1
    total = sum(v[0] for v in tot.values())
    ours = sum(v[0] for k, v in tot.items() if not k.startswith("torch:"))
    print("| share | time (ms) | launches | kernel |")
    print("|---:|---:|---:|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        if v[0] / total < 0.002:
            continue
        print("| %.2f%% | %.2f | %d | `%s` |" % (100 * v[0] / total, v[0] / 1e6, v[1], k))
    print()
    print("total %.1f ms over %d launches; hand-written sm_100a kernels: %.1f%% of device time" % (
        total / 1e6, sum(v[1] for v in tot.values()), 100 * ours / total))


if __name__ == "__main__":
    main(sys.argv[1])
