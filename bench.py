#!/usr/bin/env python
"""Benchmark of the Swapping-Autoencoder training hot path on B200 (contract: see the task brief / DESIGN.md §7).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU (native-PyTorch) path = oracle port

A "step" is one ``train_one_step`` call — a discriminator OR a generator half-step over one batch, exactly the unit
the reference's own iteration counter advances by (reference util/iter_counter.py:54); lazy R1
runs inside every 16th discriminator step.  Metric: training images / second at 256x256 with the reference's
default networks (BASELINE.json), per-GPU batch 32 (weak scaling).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RES = 256
PER_GPU_BATCH = 32
CPU_SAMPLE_BATCH = 2


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1425.5), d.get("hbm_gbs", 6480.8), "measured"
    return 1400.0, 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------------------------
def cpu_step_rate(steps, warmup, batch=CPU_SAMPLE_BATCH, threads=None):
    """Reference CPU path (oracle port of the native-PyTorch fallback): D/G half-steps at 256x256, default nets."""
    from oracle import sae_oracle as O
    from swapping_autoencoder_pytorch_b200 import default_options
    # all host cores up to 32: beyond that the reference's many small ATen ops only lose time to oversubscription
    threads = threads or min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    opt = default_options(num_gpus=0, batch_size=batch, crop_size=RES)
    model = O.OracleModel(opt, O.init_state_dict(opt, seed=0))
    trainer = O.OracleTrainer(model)
    real = torch.randn(batch, 3, RES, RES, generator=torch.Generator().manual_seed(0)).clamp(-1, 1)
    for _ in range(warmup):
        trainer.train_one_step(real)
    t0 = time.perf_counter()
    for _ in range(steps):
        trainer.train_one_step(real)
    dt = time.perf_counter() - t0
    return steps * batch / dt, dt / max(steps, 1), threads


def run_reference(args, rank):
    if rank != 0:
        return
    value, sec_per_step, threads = cpu_step_rate(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "256x256 default E/G/D/Dpatch D/G half-steps (CPU sample)", "resolution": RES,
                   "sample_batch": CPU_SAMPLE_BATCH},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "%d half-steps (D,G alternating) of batch %d at %dx%d, default nets, fp32, oracle port of the "
                                   "reference's native-PyTorch path" % (args.steps, CPU_SAMPLE_BATCH, RES, RES)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def conv_roofline(trainer, images, device):
    """FLOP-weighted throughput of the dominant kernel family (implicit-GEMM conv fprop/dgrad/wgrad launches) over
    one D + one G half-step, timed per launch with CUDA events on the launching stream."""
    from swapping_autoencoder_pytorch_b200 import backend
    k = backend.kernels()
    records = []
    orig = {n: getattr(k, n) for n in ("conv_fprop", "conv_dgrad", "conv_wgrad")}

    def wrap(name, direction):
        fn = orig[name]

        def timed(a, b, g, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(a, b, g, **kw)
            e1.record()
            flops = 2.0 * g.N * g.P * g.Q * g.K * g.R * g.S * g.C
            records.append((e0, e1, flops, k.conv_impl_for(g, direction), g.key(), name))
            return out
        return timed
    try:
        for i, n in enumerate(("conv_fprop", "conv_dgrad", "conv_wgrad")):
            setattr(k, n, wrap(n, i))
        trainer.train_one_step({"real_A": images}, 0)
        trainer.train_one_step({"real_A": images}, 0)
        torch.cuda.synchronize(device)
    finally:
        for n, fn in orig.items():
            setattr(k, n, fn)
    tot = {1: [0.0, 0.0, 0], 2: [0.0, 0.0, 0]}
    per_shape = {}
    for e0, e1, flops, impl, key, name in records:
        t = tot.setdefault(impl, [0.0, 0.0, 0])
        sec = e0.elapsed_time(e1) * 1e-3
        t[0] += flops
        t[1] += sec
        t[2] += 1
        ps = per_shape.setdefault((name, impl) + key, [0.0, 0.0, 0])
        ps[0] += flops
        ps[1] += sec
        ps[2] += 1
    dump = os.environ.get("SAE_BENCH_CONV_TABLE")
    if dump:
        with open(dump, "w") as f:
            f.write("dir impl N H W C K R S P Q stride pad_t pad_l | calls ms TFLOP/s\n")
            for kk, v in sorted(per_shape.items(), key=lambda kv: -kv[1][1]):
                f.write("%s %s | %d %.3f %.1f\n" % (kk[0], " ".join(str(x) for x in kk[1:]), v[2], v[1] * 1e3,
                                                  v[0] / max(v[1], 1e-12) / 1e12))
    return tot


def run_ours(args, rank, world, local):
    import swapping_autoencoder_pytorch_b200 as S
    from swapping_autoencoder_pytorch_b200 import _lib, backend
    import torch.distributed as dist

    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    backend.kernels()          # loads libsae_b200.so; raises if it is missing (no fallback)
    PER_GPU_BATCH = args.per_gpu_batch
    # each half-step runs as a CUDA graph replay (graphs.py) unless SAE_CUDA_GRAPHS=0
    use_graphs = os.environ.get("SAE_CUDA_GRAPHS", "1") != "0"
    opt = S.default_options(num_gpus=1, batch_size=PER_GPU_BATCH * world, crop_size=RES, cuda_graphs=use_graphs)
    torch.manual_seed(0)
    model = S.create_model(opt)
    trainer = S.create_optimizer(opt, model)

    gen = torch.Generator().manual_seed(1234 + rank)
    host = torch.randn(PER_GPU_BATCH, 3, RES, RES, generator=gen).clamp(-1, 1).pin_memory()
    resident = host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    r1_seen = []

    def timed_loop(fetch):
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        n0 = _lib.launch_count() + (trainer.graphs.replayed_launches if trainer.graphs else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r1_count = 0
        e0.record()
        for _ in range(args.steps):
            out = trainer.train_one_step({"real_A": fetch()}, 0)     # to_numpy inside reads the losses back (D2H)
            r1_count += int("D_R1" in out)
        e1.record()
        barrier()
        r1_seen.append(r1_count)
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() + (trainer.graphs.replayed_launches if trainer.graphs else 0) - n0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    # at least 4 untimed half-steps: D, G, D, G — both step kinds twice, so the caching allocator and every kernel
    # variant have reached steady state before the timed region
    n_warm = max(args.warmup, int(os.environ.get("SAE_BENCH_MIN_WARM", "4")))      # (profiling runs under ncu lower it)
    for _ in range(n_warm):
        trainer.train_one_step({"real_A": resident}, 0)
    if trainer.graphs is not None:
        trainer.graphs.warm_up(resident)          # capture the D, G and R1 graphs before the timed region
    if (trainer.graphs is None or trainer.graphs.disabled) and os.environ.get("SAE_BENCH_SKIP_R1_WARM") != "1":
        # eager execution: the lazy-R1 evaluation (every 16th discriminator step) must not meet the caching allocator's
        # cold start inside the timed region — keep stepping until one R1 has run in its natural place (<= 32 half-steps)
        for _ in range(2 * opt.R1_once_every):
            if "D_R1" in trainer.train_one_step({"real_A": resident}, 0):
                break
        trainer.train_one_step({"real_A": resident}, 0)      # the G half-step that completes the pair
        torch.cuda.synchronize(device)
    ms_dev, launches, clocks = timed_loop(lambda: resident)
    ms_e2e, _, _ = timed_loop(lambda: host.to(device, non_blocking=True))
    images = args.steps * PER_GPU_BATCH * world

    if trainer.graphs is not None:
        trainer.graphs.enabled = False            # the per-launch instrumentation pass needs eager launches ...
        for _ in range(2):                        # ... and a warm caching allocator under them
            trainer.train_one_step({"real_A": resident}, 0)
        torch.cuda.synchronize(device)
    roof = conv_roofline(trainer, resident, device)
    graph_state = "off"
    if trainer.graphs is not None:
        g = trainer.graphs
        graph_state = ("off (capture failed: %s)" % g.disabled) if g.disabled else \
            "on (%s captured; forward+backward%s per replay)" % (
                "/".join(sorted(k[0] for k in g.captured)), "+Adam" if world == 1 else "; all-reduce and Adam eager")
    peak_tf, peak_bw, peak_src = measured_peaks()
    dom = 2 if roof.get(2, [0, 0, 0])[2] > 0 else 1
    fl, sec, cnt = roof[dom]
    achieved = fl / sec / 1e12 if sec > 0 else 0.0
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        traffic_note = "%s: dram read+write per launch from %s; algorithmic bytes %d" % (tj["kernel"], tj["source"],
                                                                                         tj["algorithmic_bytes"])
    tf32_nominal = 1125.0      # dense kind::tf32 rate = half the nominal 2250 TFLOP/s bf16 rate (B200_PROFILING.md)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_tf32_nominal": tf32_nominal, "frac_of_tf32_nominal": achieved / tf32_nominal,
                "kernel": "conv implicit GEMM (%s)" % ("tcgen05+TMA TF32" if dom == 2 else "mma.sync TF32 generic"),
                "launches_timed": cnt, "peak_source": peak_src + " bf16 dense sustained; kind::tf32 peaks at half of it",
                "by_impl": {("tcgen05" if i == 2 else "generic"): {"tflops": (v[0] / v[1] / 1e12 if v[1] > 0 else 0.0),
                                                                  "share_of_flops": v[0] / max(sum(x[0] for x in roof.values()), 1.0),
                                                                  "launches": v[2]} for i, v in roof.items()}}

    if rank != 0:
        return
    # bounded CPU sample (one D + one G half-step on 2 images); N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, _, threads = cpu_step_rate(2, 0)
        cpu = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": "1 D + 1 G half-step, batch %d at %dx%d, default nets, fp32 oracle port of the reference's "
                         "native-PyTorch CPU path" % (CPU_SAMPLE_BATCH, RES, RES)}
    loss_bytes = 8 * 4
    line = {
        "metric": "training images/sec", "value": images / (ms_dev * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
        "config": {"workload": "256x256 default E/G/D/Dpatch, alternating D/G half-steps with lazy R1 (BASELINE configs[1] "
                               "shape at the metric's bs=32)", "resolution": RES, "per_gpu_batch": PER_GPU_BATCH,
                   "global_batch": PER_GPU_BATCH * world, "parallelism": "dp%d" % world,
                   "r1_evaluations_in_timed_region": r1_seen[0], "r1_once_every": opt.R1_once_every,
                   "cuda_graphs": graph_state,
                   "l2": "activations per step exceed the 126 MB L2 by >100x; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": images / (ms_e2e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": loss_bytes},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # default 32 half-steps = 16 discriminator steps: exactly one lazy-R1 evaluation falls inside the timed region
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-gpu-batch", type=int, default=PER_GPU_BATCH,
                    help="experiments only (host-overhead probes); the reported metric is defined at the default 32")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    from swapping_autoencoder_pytorch_b200.parallel import init_distributed
    rank, world, local = init_distributed("nccl")
    try:
        run_ours(args, rank, world, local)
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
