#!/usr/bin/env python
"""Benchmark of the sparse-voxel hot path: MinkUNet-34 cr1.0 forward+backward, scans/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one training step (AMP fp16 forward + backward + SGD update, as
train.py:340-372 of the reference) of MinkUNet-34 cr1.0 over one batch of B synthetic
SemanticKITTI-shaped scans (64 x 1875 rays, 0.05 m voxels, ~95 k voxels per scan) per GPU.
Prints ONE JSON line (see README / DESIGN.md for the fields):
  value     scans/s over all GPUs, inputs resident in HBM when the timed region starts
  e2e       the same through the public API with pinned HOST buffers: H2D of the batch and
            D2H of the loss inside the timed region
  roofline  the dominant kernel family (conv gather-GEMM or wgrad) timed live with CUDA events around
            every launch, in a repeat of the same K steps after the `value` region
  cpu_baseline  the reference's own CPU backend (oracle/_ref) on a bounded sub-scan
``--impl reference`` times that CPU path alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "scans/s MinkUNet-34 fwd+bwd @~120k pts/0.05 m voxel"
FULL_AZIMUTH = 1875


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("B2S_BENCH_BATCH", 16)),
                    help="scans per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--sync-bn", action="store_true")
    ap.add_argument("--pool", type=int, default=4, help="distinct batches cycled per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(budget_s: float, steps: int, warmup: int, full_voxels: int | None = None):
    """Time the reference's CPU implementation (oracle/_ref, else the oracle port) of the
    same step on a bounded sub-scan; returns (scans_per_s, info dict)."""
    from oracle.cpu_minkunet import CpuMinkUNet, kind
    from openpcseg_b200.segmentors import MinkUNet, minkunet34_config
    from openpcseg_b200.synthetic import make_batch
    # Thread count: measured on the 128-core B200 host (scripts/cpu_baseline_probe.py, 1/32
    # sub-scan): 8 threads 5.1 s, 32 threads 25.4 s - the reference's per-offset OpenMP gather
    # loops + small MKL GEMMs slow down when oversubscribed, so 8 is the fastest setting.
    cores = min(os.cpu_count() or 1, int(os.environ.get("B2S_CPU_THREADS", 8)))
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    torch.manual_seed(0)
    state = MinkUNet(minkunet34_config()).state_dict()
    if full_voxels is None:
        full_voxels = make_batch([0])["coords"].shape[0]

    def one(n_az, seed):
        b = make_batch([seed], n_azimuth=n_az)
        net = CpuMinkUNet(state)
        t0 = time.perf_counter()
        _, loss = net.forward(torch.from_numpy(b["coords"]), torch.from_numpy(b["feats"]),
                              torch.from_numpy(b["labels"]))
        loss.backward()
        return time.perf_counter() - t0, b["coords"].shape[0]

    # calibrate on a 1/32 sub-scan, then size the sample to the time budget
    t_cal, v_cal = one(max(FULL_AZIMUTH // 32, 8), 100)
    rate = v_cal / t_cal                                     # voxels per second, first guess
    per_step = budget_s / max(steps + warmup, 1)
    n_az = int(np.clip(FULL_AZIMUTH * (rate * per_step) / full_voxels, 24, FULL_AZIMUTH))
    for w in range(warmup):
        one(n_az, 200 + w)
    ts, vs = [], []
    for s in range(steps):
        t, v = one(n_az, 300 + s)
        ts.append(t)
        vs.append(v)
    scans = sum(vs) / full_voxels                            # fraction-of-scan units processed
    value = scans / sum(ts)
    info = {"value": value, "unit": "scans/s", "cores": cores, "kind": kind(),
            "sample": f"{steps} steps of fwd+bwd on a {n_az}/{FULL_AZIMUTH}-azimuth sub-scan "
                      f"(~{int(np.mean(vs))} of {full_voxels} voxels), fp32, scaled by voxel count"}
    return value, info, sum(ts) / max(len(ts), 1) * 1e3


# ------------------------------------------------------------------------- helpers
class ConvProfiler:
    """CUDA events around every conv kernel launch of the timed region (events are pre-created so
    that recording them costs ~1 us of host time per launch)."""

    def __init__(self, n_events=0):
        self.rows = []
        self.pool = [torch.cuda.Event(enable_timing=True) for _ in range(n_events)]
        self.next = 0

    def event(self):
        if self.next < len(self.pool):
            ev = self.pool[self.next]
            self.next += 1
            return ev
        return torch.cuda.Event(enable_timing=True)

    def record(self, kind, meta, start, end):
        self.rows.append((kind, meta, start, end))

    def summarise(self):
        out = {}
        totals = {}
        for kind, meta, s, e in self.rows:
            ms = s.elapsed_time(e)
            pairs = meta["pairs"]
            m = int(pairs.item()) if pairs is not None else int(meta["rows"])
            if meta["k"] == 1:
                m = int(meta["rows"])
            flops = 2.0 * m * meta["c_in"] * meta["c_out"]
            t = totals.setdefault(kind, [0.0, 0.0, 0])
            t[0] += ms
            t[1] += flops
            t[2] += 1
        for k, (ms, fl, n) in totals.items():
            out[k] = {"ms": ms, "gflop": fl / 1e9, "launches": n,
                      "tflops": (fl / 1e12) / (ms / 1e3) if ms > 0 else 0.0}
        return out


def clocks_sampler():
    try:
        return subprocess.Popen(
            ["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100",
             "-i", str(torch.cuda.current_device())],
            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def clocks_summary(proc):
    if proc is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except Exception:
        proc.kill()
        out = ""
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in out.strip().splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1]))
            mx.append(float(f[2]))
        except ValueError:
            continue
        for name, val in zip(names, f[5:9]):
            if val.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm), "reasons": sorted(reasons)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return
        budget = 150.0
        value, info, ms = cpu_reference_run(budget, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "MinkUNet-34 cr1.0 fwd+bwd, synthetic SemanticKITTI ~120k pts, "
                                       "0.05 m voxel (reference CPU path, bounded sub-scan per step)"},
                "cpu_baseline": info,
                "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from openpcseg_b200 import backend as B
    import openpcseg_b200.torchsparse as ts
    from openpcseg_b200.segmentors import MinkUNet, minkunet34_config
    from openpcseg_b200.synthetic import make_batch

    torch.manual_seed(0)
    amp = args.dtype == "fp16"
    model = MinkUNet(minkunet34_config(sync_bn=args.sync_bn and world > 1)).to(dev)
    model.train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    # SGD nesterov, lr = 0.02 per sample (minkunet_mk34_cr10.yaml:26-35), AMP GradScaler, clip 10
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9,
                          weight_decay=1e-4, nesterov=True)
    scaler = torch.amp.GradScaler("cuda", enabled=amp)

    # ---- synthetic data: a pool of distinct batches per rank, in pinned host memory
    pool = []
    for p in range(args.pool):
        seeds = [1000 * rank + 10 * p + i for i in range(args.batch)]
        b = make_batch(seeds)
        pool.append({k: torch.from_numpy(b[k]).pin_memory() for k in ("coords", "feats", "labels")})
    vox_per_scan = float(np.mean([p["coords"].shape[0] for p in pool])) / args.batch
    h2d = int(np.mean([sum(t.numel() * t.element_size() for t in p.values()) for p in pool]))
    resident = [{k: v.to(dev) for k, v in p.items()} for p in pool]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step(batch):
        lidar = ts.SparseTensor(batch["feats"], batch["coords"], 1)
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            out = net({"lidar": lidar, "targets": batch["labels"]})
        loss = out["loss"]
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        scaler.step(opt)
        scaler.update()
        return loss

    def run(n_steps, from_host, timed):
        """Returns (ms_total over the timed steps [device events], last loss value)."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        last = None
        ev0.record()
        for i in range(n_steps):
            flush_buf.zero_()                              # L2 flush between iterations
            if from_host:
                src = pool[i % len(pool)]
                batch = {k: v.to(dev, non_blocking=True) for k, v in src.items()}
            else:
                batch = resident[i % len(resident)]
            loss = step(batch)
            if from_host:
                last = float(loss.item())                  # D2H of the step's result
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    warm = max(args.warmup, 3)
    run(warm, False, False)

    # ---- timed region 1: device-resident inputs (this is `value`)
    B.STATS["launches"] = 0
    sampler = clocks_sampler() if rank == 0 else None
    ms_dev, _ = run(args.steps, False, True)
    launches = B.STATS["launches"]
    clocks = clocks_summary(sampler) if rank == 0 else None

    # ---- the same steps again with every conv launch bracketed by CUDA events (roofline numbers only:
    # the ~250 event pairs per step cost a few % of throughput, so they stay out of `value`)
    prof = ConvProfiler(n_events=args.steps * 420)          # ~190 conv launches x 2 events per step
    B.PROFILER = prof
    ms_prof, _ = run(args.steps, False, True)
    B.PROFILER = None
    conv = prof.summarise()

    # ---- timed region 2: end to end from pinned host buffers
    run(2, True, False)
    ms_e2e, last_loss = run(args.steps, True, True)

    scans = args.batch * args.steps * world
    value = scans / (ms_dev / 1e3)
    e2e = scans / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    # The dominant kernel: forward and input-gradient are the SAME kernel (gather_gemm_tc3_kernel,
    # different maps / weight transposes), the weight gradient is a second kernel (wgrad_tc_kernel).
    roof = None
    if conv:
        fams = {"conv gather-GEMM (fwd+dgrad), gather_gemm_tc3_kernel": ("fwd", "dgrad"),
                "conv wgrad, wgrad_tc_kernel": ("wgrad",)}
        tot = {name: sum(conv[k]["ms"] for k in ks if k in conv) for name, ks in fams.items()}
        name = max(tot, key=tot.get)
        ks = [k for k in fams[name] if k in conv]
        fam_ms = sum(conv[k]["ms"] for k in ks)
        fam_fl = sum(conv[k]["gflop"] for k in ks)
        achieved = (fam_fl / 1e3) / (fam_ms / 1e3)
        roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS, MEASURED_PEAKS.json)",
                "traffic": None, "launches_timed": sum(conv[k]["launches"] for k in ks),
                "share_of_step": fam_ms / ms_prof,
                "per_family": {k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in conv.items()},
                "note": "achieved = useful FLOPs 2*M*Cin*Cout (M = kernel-map pairs) / CUDA-event time, "
                        "summed over every launch in the timed region; DRAM traffic per launch is layer "
                        "dependent - see profiles/ (ncu --set full captures) for representative layers"}

    cpu_info = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            _, cpu_info, _ = cpu_reference_run(args.cpu_budget_s, 1, 0, int(vox_per_scan))
        except Exception as exc:                               # never lose the GPU line
            cpu_info = {"value": None, "unit": "scans/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"failed: {exc}"}

    line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if amp else "f32", "data": "synthetic",
            "config": {"workload": "MinkUNet-34 cr1.0 fwd+bwd(+SGD step), synthetic SemanticKITTI "
                                   "64x1875 rays, 0.05 m voxel",
                       "scans_per_gpu_per_step": args.batch, "voxels_per_scan": int(vox_per_scan),
                       "amp": amp, "sync_bn": bool(args.sync_bn and world > 1),
                       "parallelism": f"dp{world}", "l2": "256 MiB flush write before every step",
                       "peaks": pk},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu_info}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
