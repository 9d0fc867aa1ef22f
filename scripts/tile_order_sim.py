#!/usr/bin/env python
"""CPU simulation of the conv kernels' tile composition: how many (tile, offset) steps does a row order
cost?  (The step count is what the gather-GEMM time follows, profiles/r1_conv_ablation.txt.)

usage: python scripts/tile_order_sim.py [--batch 2] [--stride 1]
"""
import argparse
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openpcseg_b200.synthetic import make_batch          # noqa: E402


def offsets3(stride):
    r = np.array([-1, 0, 1]) * stride
    return np.array([[x, y, z] for z in r for y in r for x in r])      # x fastest (odd kernels)


def pack(c):
    c = c.astype(np.int64)
    return ((c[:, 3] << 54) | ((c[:, 0] + 131072) << 36) | ((c[:, 1] + 131072) << 18) | (c[:, 2] + 131072))


def neighbour_bits(coords, stride):
    key = pack(coords)
    order = np.argsort(key)
    skey = key[order]
    bits = np.zeros(coords.shape[0], dtype=np.int64)
    sizes = []
    for k, off in enumerate(offsets3(stride)):
        q = coords.copy()
        q[:, :3] += off
        qk = pack(q)
        pos = np.searchsorted(skey, qk)
        pos[pos >= skey.shape[0]] = 0
        hit = skey[pos] == qk
        bits |= hit.astype(np.int64) << k
        sizes.append(int(hit.sum()))
    return bits, np.array(sizes)


def steps_of(bits_sorted):
    n = bits_sorted.shape[0]
    pad = (-n) % 128
    b = np.concatenate([bits_sorted, np.zeros(pad, dtype=np.int64)]).reshape(-1, 128)
    union = np.bitwise_or.reduce(b, axis=1)
    pop = np.zeros(union.shape[0], dtype=np.int64)
    for k in range(27):
        pop += (union >> k) & 1
    return int(pop.sum()), pop


def remap_bits(bits, sizes, rarest_msb=True):
    rank = np.argsort(np.argsort(sizes, kind="stable"), kind="stable")      # 0 = rarest
    out = np.zeros_like(bits)
    for k in range(27):
        pos = 26 - rank[k] if rarest_msb else rank[k]
        out |= ((bits >> k) & 1) << pos
    return out


def coarse(coords, shift):
    c = coords.astype(np.int64)
    return (((c[:, 2] >> shift) & 0x3FF) << 26) | (((c[:, 0] >> shift) & 0x1FFF) << 13) | ((c[:, 1] >> shift) & 0x1FFF)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--stride", type=int, default=1)
    a = ap.parse_args()
    b = make_batch(list(range(a.batch)))
    coords = b["coords"].astype(np.int64)
    s = a.stride
    while s > 1 and False:
        pass
    if a.stride > 1:
        c = coords.copy()
        c[:, :3] = (c[:, :3] // a.stride) * a.stride
        coords = np.unique(c, axis=0)
    shift = max(int(a.stride).bit_length() - 1, 0)
    bits, sizes = neighbour_bits(coords, a.stride)
    n, m = coords.shape[0], int(sizes.sum())
    tiles = (n + 127) // 128
    print(f"N={n} M={m} tiles={tiles} lower bound steps={m / 128:.0f} ({m / 128 / tiles:.2f}/tile)")

    def report(name, order):
        st, pop = steps_of(bits[order])
        print(f"{name:44s} steps {st:7d}  {st / tiles:5.2f}/tile  active {st / (27.0 * tiles):.3f}  fill {m / (st * 128.0):.3f}")
        return st

    rng = np.random.default_rng(0)
    report("random (hash order)", rng.permutation(n))
    report("(b,z,x,y)", np.argsort((coords[:, 3] << 54) | (coords[:, 2] << 36) | (coords[:, 0] << 18) | coords[:, 1]))
    rb = remap_bits(bits, sizes)
    cur = np.argsort((rb << 36) | coarse(coords, shift), kind="stable")
    report("current: pattern (rarest msb) | coarse zxy", cur)
    report("pattern (commonest msb) | coarse zxy", np.argsort((remap_bits(bits, sizes, False) << 36) | coarse(coords, shift), kind="stable"))
    # popcount first, then pattern
    pc = np.zeros(n, dtype=np.int64)
    for k in range(27):
        pc += (bits >> k) & 1
    report("popcount | pattern | coarse", np.argsort((pc << 58) | (rb << 31) | (coarse(coords, shift) >> 5), kind="stable"))
    # drop the commonest bits from the key (they are nearly everywhere): pattern of the 18 rarest only
    for keep in (9, 12, 15, 18, 21):
        mask_bits = rb >> (27 - keep)
        report(f"{keep} rarest bits | coarse zxy", np.argsort((mask_bits << 36) | coarse(coords, shift), kind="stable"))
    return bits, sizes, coords


if __name__ == "__main__":
    main()
