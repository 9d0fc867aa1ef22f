#!/usr/bin/env python
"""Where the GPU idles inside one training step: kernel timeline (torch.profiler, CPU + CUDA activities, chrome
trace) -> busy time, idle gaps between consecutive device activities, the gaps that contain the end of a host
sync (cudaStreamSynchronize / cudaMemcpy D2H), and the (previous kernel -> next kernel) pairs with the most idle.
nsys is not in this image; this is the timeline evidence for DESIGN.md "host syncs" and the N = 8 drift.
    python scripts/step_timeline.py [--batch 16] [--steps 3]"""
import argparse
import collections
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--pool", type=int, default=3)
    a = ap.parse_args()
    import openpcseg_b200.torchsparse as ts
    from openpcseg_b200.segmentors import MinkUNet, minkunet34_config
    from openpcseg_b200.synthetic import make_model_batch
    dev = torch.device("cuda", 0)
    model = MinkUNet(minkunet34_config()).to(dev).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4, fused=True)
    scaler = torch.amp.GradScaler("cuda")
    pool = []
    for p in range(a.pool):
        arr = make_model_batch("voxel", list(range(p * a.batch, (p + 1) * a.batch)))
        pool.append({k: torch.from_numpy(v).to(dev) for k, v in arr.items() if isinstance(v, np.ndarray)})

    def step(i):
        res = pool[i % len(pool)]
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            loss = model({"lidar": ts.SparseTensor(res["feats"], res["coords"], 1), "targets": res["labels"]})["loss"]
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()

    for i in range(3 * a.pool):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(a.steps):
            step(i)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "b2s_step_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    gpu = sorted(((e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev
                  if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda r: r[0])
    syncs = sorted((e["ts"] + e["dur"], e["name"], e["dur"]) for e in ev
                   if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver")
                   and ("Synchronize" in e["name"] or "cudaMemcpy" in e["name"]) and e["dur"] > 20)
    t0, t1 = gpu[0][0], max(g[1] for g in gpu)
    busy, gaps, end = 0.0, [], gpu[0][0]
    prev = gpu[0][2]
    for s, e, name in gpu:
        if s > end:
            gaps.append((s - end, end, s, prev, name))
            busy += e - s
        else:
            busy += max(e - max(s, end), 0.0)
        if e > end:
            end, prev = e, name
    span = t1 - t0
    idle = sum(g[0] for g in gaps)
    print(f"# MinkUNet-34 batch {a.batch}, {a.steps} steps under torch.profiler (CPU+CUDA): span {span / a.steps / 1e3:.2f} ms/step, "
          f"device busy {busy / a.steps / 1e3:.2f} ms/step, idle {idle / a.steps / 1e3:.2f} ms/step in {len(gaps) // a.steps} gaps/step")
    for lo, hi in ((0, 5), (5, 20), (20, 100), (100, 1000), (1000, 1e9)):
        sel = [g for g in gaps if lo <= g[0] < hi]
        print(f"  gaps {lo:>4}-{hi if hi < 1e9 else 'inf':>5} us: {len(sel) / a.steps:7.1f} per step, {sum(g[0] for g in sel) / a.steps / 1e3:6.3f} ms/step")
    # gaps in which a host sync returned (the CPU had been blocked and the queue was empty)
    sync_idle, sync_hits, si = 0.0, collections.Counter(), 0
    marked = set()
    for t_end, name, dur in syncs:
        for gi, g in enumerate(gaps):
            if g[1] - 5 <= t_end <= g[2] + 5 and gi not in marked:
                marked.add(gi)
                sync_idle += g[0]
                sync_hits[name] += 1
                break
    print(f"  host syncs > 20 us per step: {len(syncs) / a.steps:.1f}; idle in the gaps where one returned: "
          f"{sync_idle / a.steps / 1e3:.3f} ms/step  {dict(sync_hits)}")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for d, _, _, p, n in gaps:
        k = (p[:48], n[:48])
        agg[k][0] += d
        agg[k][1] += 1
    print("  most idle by (previous activity -> next activity):")
    for (p, n), (d, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"    {d / a.steps / 1e3:6.3f} ms/step {c / a.steps:6.1f}x  {p}  ->  {n}")
    # position of the idle time inside the step (tenths of the span of each step)
    per = span / a.steps
    hist = np.zeros(10)
    for d, s, _, _, _ in gaps:
        hist[min(int(((s - t0) % per) / per * 10), 9)] += d
    print("  idle by tenth of the step (ms/step): " + " ".join(f"{h / a.steps / 1e3:.2f}" for h in hist))


if __name__ == "__main__":
    main()
