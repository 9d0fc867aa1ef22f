#!/usr/bin/env python
"""Per-kernel CUDA-time breakdown (torch.profiler) of one training step of a bench config.
    python scripts/profile_models.py --config rpvnet34 [--model-src reference|native] [--batch B] [--top 30]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="minkunet34")
    ap.add_argument("--model-src", default=None)
    ap.add_argument("--backend", default="b2s")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    kind, dbatch, _, _ = BN.CONFIGS[a.config]
    batch = a.batch or dbatch
    src = a.model_src or ("native" if a.config == "minkunet34" else "reference")
    from openpcseg_b200.synthetic import make_model_batch
    dev = torch.device("cuda", 0)
    ns = None
    if src == "native":
        import openpcseg_b200.torchsparse as ts
        from openpcseg_b200.segmentors import MinkUNet, minkunet34_config
        model = MinkUNet(minkunet34_config()).to(dev)
    else:
        from baseline import loader
        ns = loader.activate(a.backend)
        model = ns.build_model(a.config).to(dev)
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    arrays = make_model_batch(kind, list(range(batch)))
    res = {k: torch.from_numpy(v).to(dev) for k, v in arrays.items() if isinstance(v, np.ndarray)}

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            if ns is None:
                loss = model({"lidar": ts.SparseTensor(res["feats"], res["coords"], 1), "targets": res["labels"]})["loss"]
            else:
                loss = model(ns.batch_dict(res, dev))[0]["loss"]
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        scaler.step(opt)
        scaler.update()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / a.steps / 1e3, e.count // a.steps) for e in prof.key_averages()
            if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    print(f"# {a.config} ({src} on {a.backend}), batch {batch}: {total:.2f} ms of kernels per step, "
          f"{sum(r[2] for r in rows)} launches per step")
    for name, ms, n in rows[:a.top]:
        print(f"{ms:9.3f} ms {100 * ms / total:5.1f}% {n:5d}  {name[:110]}")


if __name__ == "__main__":
    main()
