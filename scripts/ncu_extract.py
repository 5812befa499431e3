#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (from `ncu --set full`), as text for profiles/.
usage: python scripts/ncu_extract.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__cluster_size", "launch__waves_per_multiprocessor",
    "smsp__cycles_active.avg", "sm__inst_executed.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel: %s" % d["Kernel Name"])
        print("  grid %s block %s" % (d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                print("  %-96s %s %s" % (k, d[k], units[hdr.index(k)]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
