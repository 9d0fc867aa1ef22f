#!/usr/bin/env python
"""Achieved HBM GB/s of the memory-bound stages of the path (hash, table, kernel map, point<->voxel, batch norm,
range ops) at BASELINE scale: CUDA-event time (median, L2 flushed between runs) and the ALGORITHMIC bytes of
SURVEY.md 8(d) / DESIGN.md 3 per launch, against the measured copy bandwidth in MEASURED_PEAKS.json.

    python scripts/membound_ops.py [--batch 4] [--once]        (--once: one call per op, for an ncu capture)
ncu (per-launch DRAM bytes; see profiles/README.md):
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/membound_ncu.csv python scripts/membound_ops.py --once
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--iters", type=int, default=7)
    a = ap.parse_args()
    from openpcseg_b200 import backend as B
    import openpcseg_b200.torchsparse as ts
    from openpcseg_b200.synthetic import make_batch, make_model_batch
    F = ts.nn.functional
    dev = torch.device("cuda", 0)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    b = make_batch(list(range(a.batch)))
    coords = torch.from_numpy(b["coords"]).to(dev)
    n = coords.shape[0]
    order = torch.argsort(F.sphash(coords))                      # the model's row order (ascending hash)
    coords = coords[order].contiguous()
    pts = coords.float().contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def run(name, fn, alg_bytes, note=""):
        if a.once:
            fn()
            return
        for _ in range(2):
            fn()
        ts_ms = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts_ms.append(e0.elapsed_time(e1))
        ms = sorted(ts_ms)[len(ts_ms) // 2]
        gbs = alg_bytes / ms / 1e6
        rows.append((name, ms * 1e3, alg_bytes / 1e6, gbs, gbs / peak, note))

    offs = ts.nn.utils.get_kernel_offsets(3, 1, 1, dev)
    k = offs.shape[0]
    h = F.sphash(coords)
    run("hash_kernel (sphash)", lambda: B.hash_coords(coords), 16 * n + 8 * n)
    run("kernel_hash_kernel K=27", lambda: B.kernel_hash(coords, offs), 16 * n + 8 * k * n)
    slots = int(B._lib.lib().b2s_table_slots(n))
    run("table_build (keys)", lambda: B.HashTable.from_keys(h), 8 * n + 12 * slots, "memset of the table included")
    table = B.HashTable.from_keys(h)
    q = h[torch.randperm(n, device=dev)]
    run("table_query", lambda: table.query(q), 16 * n + 32 * n, "32 B probe sector per query")
    idx32 = torch.randint(0, n, (n,), device=dev, dtype=torch.int32)
    run("count_kernel", lambda: B.count(idx32, n), 4 * n + 4 * n)
    run("kmap_probe (k3, fused hash+table+probe)", lambda: B.kmap_build(coords, coords, offs, want_nbr_in=False),
        16 * n + 12 * slots + 16 * n + 4 * k * n + 32 * k * n, "upper bound: one 32 B sector per probe")
    run("downsample k2s2 (pack+sort+unique)", lambda: B.downsample_coords(coords, (2, 2, 2), (2, 2, 2), (1, 1, 1)),
        16 * n + 5 * 16 * n, "5 radix passes over 8 B keys, read+write; one host sync for the count")
    nbr, _, nbsizes, _, _ = B.kmap_build(coords, coords, offs, want_nbr_in=False)
    run("tile_order_key", lambda: B.tile_order_key(nbr, nbsizes, coords, 0), 4 * k * n + 16 * n + 12 * n)
    keys, bits = B.tile_order_key(nbr, nbsizes, coords, 0)
    perm = torch.argsort(keys).int()
    run("tile_steps (mask+scan+fill, TR=128)", lambda: B.tile_steps(nbr, perm, bits, 128),
        8 * n + 4 * n + 2 * int(0.25 * 4 * k * n), "active quarter of the map read (32 B sectors) and written")
    for c in (4, 96):
        for dt in (torch.float32, torch.float16):
            e = 4 if dt == torch.float32 else 2
            feats = torch.randn(n, c, device=dev).to(dt)
            vidx = torch.arange(n, device=dev, dtype=torch.int32)
            cnt = torch.ones(n, device=dev, dtype=torch.int32)
            tag = f"C={c} {'fp32' if e == 4 else 'fp16'}"
            run(f"voxelize_fwd {tag}", lambda: B.voxelize_forward(feats, vidx, cnt), (e * c + 4) * n + 4 * n + e * c * n)
            run(f"voxelize_bwd {tag}", lambda: B.voxelize_backward(feats, vidx, cnt, n), (e * c + 8) * n + e * c * n)
    for stride, c in ((1, 96), (4, 128), (16, 256)):
        lv = coords
        for s in range(int(np.log2(stride))):
            lv = B.downsample_coords(lv, (2, 2, 2), (2, 2, 2), (2 ** s,) * 3)
        nv = lv.shape[0]
        run(f"trilinear_map stride {stride}", lambda: B.trilinear_map(pts, lv, stride), 16 * n + 32 * 8 * n + 64 * n,
            "8 probes per point, one sector each")
        tidx, tw = B.trilinear_map(pts, lv, stride)
        vf = torch.randn(nv, c, device=dev).half()
        gp = torch.randn(n, c, device=dev).half()
        corners = 1 if stride == 1 else 8
        run(f"devoxelize_fwd stride {stride} C={c} fp16", lambda: B.devoxelize_forward(vf, tidx, tw),
            64 * n + 2 * c * corners * n + 2 * c * n, f"{corners} corner rows per point (L2 hits mostly)")
        run(f"devoxelize_bwd stride {stride} C={c} fp16", lambda: B.devoxelize_backward(gp, tidx, tw, nv),
            64 * n + 2 * c * n + 4 * c * corners * n + 6 * c * nv, "fp32 reds into the scratch + conversion pass")
    for c in (32, 96, 256):
        x = torch.randn(n, c, device=dev).half()
        res = torch.randn(n, c, device=dev).half()
        gm, bt = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        run(f"bn_forward(+res,+relu) C={c} fp16 (stats+finalize+apply)",
            lambda: B.bn_forward(x, res, gm, bt, rm, rv, 1e-5, 0.1, True), 2 * c * n + 3 * 2 * c * n)
        y, mean, invstd, _ = B.bn_forward(x, res, gm, bt, rm, rv, 1e-5, 0.1, True)
        dy = torch.randn(n, c, device=dev).half()
        run(f"bn_backward(+res,+relu) C={c} fp16 (reduce+apply)",
            lambda: B.bn_backward(dy, y, x, mean, invstd, gm, True, True), 3 * 2 * c * n + 5 * 2 * c * n)
        y2, mean2, invstd2, ss2 = B.bn_forward(x, None, gm, bt, rm, rv, 1e-5, 0.1, True)
        run(f"bn_backward(+relu, mask from x) C={c} fp16 (reduce+apply)",
            lambda: B.bn_backward(dy, None, x, mean2, invstd2, gm, True, False, ss2), 2 * 2 * c * n + 3 * 2 * c * n)
    f256 = torch.randn(n, 256, device=dev)
    vi = torch.randint(0, n // 4, (n,), device=dev)
    run("scatter_max C=256 fp32 (Cylinder3D)", lambda: B.scatter_max(f256, vi, n // 4), 4 * 256 * n + 8 * n + 12 * 256 * (n // 4),
        "key pass + argmin pass + decode")
    fb = make_model_batch("fusion", list(range(a.batch)))
    pxpy = torch.from_numpy(fb["range_pxpy"]).to(dev)
    ip = torch.cat([pxpy[:, :1], (pxpy[:, 1:] + 1) / 2 * torch.tensor([2047.0, 63.0], device=dev)], 1).int().contiguous()
    npx = ip.shape[0]
    run("map_count 64x2048", lambda: B.map_count(ip, a.batch, 64, 2048), 12 * npx + 4 * a.batch * 64 * 2048)
    cm = B.map_count(ip, a.batch, 64, 2048)
    pf = torch.randn(npx, 56, device=dev)
    run("denselize_fwd C=56", lambda: B.denselize_forward(pf, cm, ip), (4 * 56 + 16) * npx + 4 * 56 * a.batch * 64 * 2048,
        "memset of the image included")
    gd = torch.randn(a.batch, 56, 64, 2048, device=dev)
    run("denselize_bwd C=56", lambda: B.denselize_backward(gd, cm, ip), (4 * 56 + 16) * npx + 4 * 56 * npx)
    torch.cuda.synchronize()
    if a.once:
        return
    print(f"# batch {a.batch}: {n} voxels; HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json); times = median of {a.iters}, "
          f"L2 flushed, incl. the op's memsets and helper launches")
    print(f"{'op':58s} {'us':>9s} {'alg MB':>9s} {'GB/s':>8s} {'of peak':>8s}  note")
    for name, us, mb, gbs, frac, note in rows:
        print(f"{name:58s} {us:9.1f} {mb:9.1f} {gbs:8.0f} {100 * frac:7.1f}%  {note}")


if __name__ == "__main__":
    main()
