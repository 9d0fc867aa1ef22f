"""Per-layer timing of the sparse-conv kernels on real kernel maps of synthetic scans.

python scripts/conv_microbench.py [--batch B] [--iters N] [--layers name,...] [--once]
Reports, per (level, C_in->C_out) and per kernel (fwd / dgrad / wgrad): time, useful TFLOP/s
(2*M*Cin*Cout), dense-equivalent TFLOP/s (2*K*N*Cin*Cout) and algorithmic GB/s.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import openpcseg_b200.torchsparse as ts  # noqa: E402
from openpcseg_b200 import backend as B  # noqa: E402
from openpcseg_b200.synthetic import make_batch  # noqa: E402

F = ts.nn.functional
LAYERS = {  # name: (level, cin, cout)
    "L0_32x32": (0, 32, 32), "L0_96x96": (0, 96, 96), "L0_128x96": (0, 128, 96),
    "L1_32x32": (1, 32, 32), "L1_96x96": (1, 96, 96),
    "L2_64x64": (2, 64, 64), "L2_128x128": (2, 128, 128), "L2_192x128": (2, 192, 128),
    "L3_128x128": (3, 128, 128), "L3_256x256": (3, 256, 256), "L3_384x256": (3, 384, 256),
    "L4_256x256": (4, 256, 256),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--layers", default=",".join(LAYERS))
    ap.add_argument("--once", action="store_true", help="one launch per kernel (for ncu)")
    ap.add_argument("--hash-order", action="store_true",
                    help="put every level in ascending-hash row order (what initial_voxelize produces)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    b = make_batch(list(range(a.batch)))
    coords = torch.from_numpy(b["coords"]).to(dev)
    levels = [coords]
    for lv in range(4):
        levels.append(F.spdownsample(levels[-1], 2, 2, 2 ** lv))
    if a.hash_order:
        levels = [c[torch.argsort(F.sphash(c))].contiguous() for c in levels]
    kmaps = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f"batch {a.batch}: voxels per level {[int(c.shape[0]) for c in levels]}")
    seen_levels = set()
    for name in a.layers.split(","):
        lv, cin, cout = LAYERS[name]
        c = levels[lv]
        if lv not in kmaps:
            kmaps[lv] = F.build_kernel_map(c, c, 3, (2 ** lv,) * 3)
        km = kmaps[lv]
        n = c.shape[0]
        pairs, total = km.pairs()
        m = int(total.item())
        x = torch.randn(n, cin, device=dev, dtype=torch.float16)
        gy = torch.randn(n, cout, device=dev, dtype=torch.float16)
        w = (torch.randn(27, cin, cout, device=dev) / (27 * cin) ** 0.5).half()
        _, okw = km.gather_args("out", x, cin, cout)
        _, gkw = km.gather_args("in", gy, cout, cin)
        wk = B.weight_to_kmajor(w)
        runs = {
            "fwd": lambda: B.conv_gather_gemm(x, w, n_rows=n, transpose_w=False, weight_kmajor=wk, **okw),
            "dgrad": lambda: B.conv_gather_gemm(gy, w, n_rows=n, transpose_w=True, **gkw),
            "wgrad": lambda: B.conv_wgrad(x, gy, 27, *km.wgrad_pairs(x), False),
        }
        omask = okw["steps"][0] if "steps" in okw else None
        if omask is not None and name[:2] not in seen_levels:
            seen_levels.add(name[:2])
            steps = sum(bin(v & 0xFFFFFFFF).count("1") for v in omask.cpu().flatten().tolist())
            tiles = omask.shape[0]
            print(f"# {name[:2]}: {tiles} row tiles, {steps} active (tile, offset) steps = "
                  f"{steps / (27.0 * tiles):.3f} of 27/tile; pairs/step {m / max(steps, 1):.1f} of 128")
        useful = 2.0 * m * cin * cout
        dense = 2.0 * 27 * n * cin * cout
        for kind, fn in runs.items():
            if a.once:
                fn()
                continue
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
            if kind == "wgrad":
                byt = 2 * (cin + cout) * m + 8 * m + 4 * 27 * cin * cout
            else:
                cr, cs = (cin, cout) if kind == "fwd" else (cout, cin)
                byt = 2 * cr * m + 2 * cs * n + 4 * 27 * n + 2 * 27 * cin * cout
            print(f"{name:12s} {kind:5s} N={n:7d} M={m:8d} {ms * 1e3:8.1f} us  useful {useful / ms / 1e9:7.1f} TF/s  "
                  f"dense-eq {dense / ms / 1e9:7.1f} TF/s  alg {byt / ms / 1e6:7.1f} GB/s")
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
