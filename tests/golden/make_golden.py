#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (build container only).

The reference (torchsparse 1.4.0 python package from
/root/reference/package/torchsparse.zip + its CPU backend compiled by
oracle/build_ref.py, + pcseg's minkunet/utils.py) is imported here, fed seeded
inputs, and its outputs are frozen as small fixtures.  Nothing of the reference
is copied into the repo; only inputs/outputs (numbers) are stored.  The fixtures
travel to the GPU box, where /root/reference does not exist.

All clouds use a single batch index per call because the reference's CPU
kernel-hash reads the batch word of point 0 for every point
(TS/backend/hash/hash_cpu.cpp:29); multi-batch behaviour follows the CUDA
kernel (hash_cuda.cu:41-46) and is covered by oracle-vs-CUDA tests instead.

Run:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import warnings
import zipfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def import_reference(tmp: str):
    from oracle import build_ref
    build_ref.build()
    backend = build_ref.load()
    assert backend is not None, "reference CPU backend not built"
    zipfile.ZipFile("/root/reference/package/torchsparse.zip").extractall(tmp)
    sys.modules["torchsparse.backend"] = backend
    sys.path.insert(0, os.path.join(tmp, "torchsparse"))
    import torchsparse  # noqa: F401
    torchsparse.backend = backend
    spec = importlib.util.spec_from_file_location(
        "ref_mink_utils", "/root/reference/pcseg/model/segmentor/voxel/minkunet/utils.py")
    mu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mu)
    return torchsparse, mu


def synth_cloud(rng, n, extent, batch_idx=0):
    """Unique int32 voxel coords [V, 4] roughly on two surfaces (so neighbours exist)."""
    xy = rng.integers(0, extent, size=(n, 2))
    z_ground = (xy[:, 0] // 7 + xy[:, 1] // 9) % 5
    z_wall = rng.integers(0, extent // 2, size=n)
    z = np.where(rng.random(n) < 0.7, z_ground, z_wall)
    c = np.unique(np.stack([xy[:, 0], xy[:, 1], z], 1), axis=0)
    rng.shuffle(c)
    b = np.full((c.shape[0], 1), batch_idx)
    return np.concatenate([c, b], 1).astype(np.int32)


def main():
    tmp = tempfile.mkdtemp(prefix="b2s_golden_")
    ts, mu = import_reference(tmp)
    import torchsparse.nn.functional as F
    from torchsparse import PointTensor, SparseTensor
    from torchsparse.nn.utils import get_kernel_offsets

    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    T = torch.from_numpy

    # ---------------------------------------------------------------- hash + offsets
    g = {}
    ka = np.array([[0, 0, 0, 0], [1, 2, 3, 0], [-1, 0, 5, 0], [1686, 1603, 66, 0], [7, 7, 7, 1],
                   [2147483647, -2147483648, 0, 3], [0, 0, 2, 0], [0, 1, 2, 0], [0, 1, 5, 0]],
                  dtype=np.int32)
    g["ka_coords"] = ka
    g["ka_hash"] = F.sphash(T(ka)).numpy()
    rc = rng.integers(-3000, 3000, size=(777, 4)).astype(np.int32)
    rc[:, 3] = 2                                   # one batch index for the whole call
    g["rand_coords"] = rc
    g["rand_hash"] = F.sphash(T(rc)).numpy()
    for name, (ks, st) in {"k3": (3, 1), "k2s4": (2, 4), "k133": ((1, 3, 3), 1),
                           "k313": ((3, 1, 3), 2), "k311": ((3, 1, 1), 1), "k3s8": (3, 8)}.items():
        off = get_kernel_offsets(ks, st)
        g[f"off_{name}"] = off.numpy()
        g[f"khash_{name}"] = F.sphash(T(rc), off).numpy()
    np.savez_compressed(os.path.join(HERE, "hash_offsets.npz"), **g)

    # -------------------------------------------------- 5-voxel known-answer
    g = {}
    c5 = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [1, 1, 1, 0], [3, 3, 3, 0]], np.int32)
    x = SparseTensor(T(np.arange(5, dtype=np.float32)[:, None].copy()), T(c5), 1)
    x.cmaps[x.stride] = x.coords
    y = F.conv3d(x, torch.ones(27, 1, 1), 3)
    nb, ns, sz = x.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    g.update(coords=c5, k3_nbmaps=nb.numpy(), k3_nbsizes=ns.numpy(), k3_out=y.feats.numpy())
    y2 = F.conv3d(x, torch.ones(8, 1, 1), 2, stride=2)
    nb, ns, sz = x.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]
    g.update(k2s2_coords=y2.coords.numpy(), k2s2_nbmaps=nb.numpy(), k2s2_nbsizes=ns.numpy(),
             k2s2_out=y2.feats.numpy())
    y3 = F.conv3d(y2, torch.ones(8, 1, 1), 2, stride=2, transposed=True)
    g.update(k2s2t_out=y3.feats.numpy(), k2s2t_coords=y3.coords.numpy())
    np.savez_compressed(os.path.join(HERE, "five_voxel.npz"), **g)

    # ------------------------------------------------ random cloud: maps + conv fwd/bwd
    g = {}
    coords = synth_cloud(rng, 2600, 48, batch_idx=0)
    g["coords"] = coords
    cin, cout = 8, 12
    feats = torch.randn(coords.shape[0], cin)
    g["feats"] = feats.numpy()

    def run_conv(x, w, ks, stride=1, transposed=False, tag=""):
        w = w.clone().requires_grad_(True)
        xin = SparseTensor(x.feats.clone().requires_grad_(True), x.coords, x.stride)
        xin.cmaps, xin.kmaps = x.cmaps, x.kmaps
        y = F.conv3d(xin, w, ks, stride=stride, transposed=transposed)
        go = torch.randn_like(y.feats)
        y.feats.backward(go)
        g[f"{tag}_w"] = w.detach().numpy()
        g[f"{tag}_out"] = y.feats.detach().numpy()
        g[f"{tag}_coords"] = y.coords.numpy()
        g[f"{tag}_gout"] = go.numpy()
        g[f"{tag}_gin"] = xin.feats.grad.numpy()
        g[f"{tag}_gw"] = w.grad.numpy()
        return y

    x0 = SparseTensor(feats, T(coords), 1)
    x0.cmaps[x0.stride] = x0.coords
    y = run_conv(x0, torch.randn(27, cin, cout) * 0.2, 3, tag="k3")
    nb, ns, sz = x0.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    g.update(k3_nbmaps=nb.numpy(), k3_nbsizes=ns.numpy())
    x1 = run_conv(x0, torch.randn(8, cin, cout) * 0.2, 2, stride=2, tag="k2s2")
    nb, ns, sz = x0.kmaps[((1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))]
    g.update(k2s2_nbmaps=nb.numpy(), k2s2_nbsizes=ns.numpy())
    x1d = SparseTensor(x1.feats.detach(), x1.coords, x1.stride)
    x1d.cmaps, x1d.kmaps = x1.cmaps, x1.kmaps
    y = run_conv(x1d, torch.randn(27, cout, cout) * 0.2, 3, tag="s2k3")      # k3 at stride 2
    nb, ns, sz = x1.kmaps[((2, 2, 2), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    g.update(s2k3_nbmaps=nb.numpy(), s2k3_nbsizes=ns.numpy())
    y = run_conv(x1d, torch.randn(8, cout, cin) * 0.2, 2, stride=2, transposed=True, tag="k2s2t")
    # asymmetric kernels (Cylinder3D) on a fresh tensor
    for tag, ks in {"k133": (1, 3, 3), "k313": (3, 1, 3), "k311": (3, 1, 1)}.items():
        xa = SparseTensor(feats, T(coords), 1)
        xa.cmaps[xa.stride] = xa.coords
        run_conv(xa, torch.randn(int(np.prod(ks)), cin, cout) * 0.2, ks, tag=tag)
        nb, ns, sz = xa.kmaps[((1, 1, 1), ks, (1, 1, 1), (1, 1, 1))]
        g.update({f"{tag}_nbmaps": nb.numpy(), f"{tag}_nbsizes": ns.numpy()})
    # slow-path downsample: k3 stride 2 and stride (2,2,1) (cylinder_ts.py:204-214)
    for tag, st in {"k3s2": (2, 2, 2), "k3s221": (2, 2, 1)}.items():
        xa = SparseTensor(feats, T(coords), 1)
        xa.cmaps[xa.stride] = xa.coords
        run_conv(xa, torch.randn(27, cin, cout) * 0.2, 3, stride=st, tag=tag)
        nb, ns, sz = xa.kmaps[((1, 1, 1), (3, 3, 3), st, (1, 1, 1))]
        g.update({f"{tag}_nbmaps": nb.numpy(), f"{tag}_nbsizes": ns.numpy()})
    np.savez_compressed(os.path.join(HERE, "conv_maps.npz"), **g)

    # --------------------------------------------------- point <-> voxel (utils.py path)
    g = {}
    npts = 3000
    pts = np.concatenate([rng.uniform(0, 40, size=(npts, 3)), np.zeros((npts, 1))], 1).astype(np.float32)
    pts[: npts // 3, :3] = np.floor(pts[: npts // 3, :3])     # some integer-valued points
    pf = rng.standard_normal((npts, 4)).astype(np.float32)
    g["pts"], g["pt_feats"] = pts, pf
    z = PointTensor(T(pf.copy()), T(pts.copy()))
    x0 = mu.initial_voxelize(z, 0.05, 0.05)
    g["iv_coords"], g["iv_feats"] = x0.coords.numpy(), x0.feats.numpy()
    g["iv_idx_query"] = z.additional_features["idx_query"][1].numpy()
    g["iv_counts"] = z.additional_features["counts"][1].numpy()
    # stride-1 devoxelize + a stride-2 level
    vf = torch.randn(x0.coords.shape[0], 6)
    xs = SparseTensor(vf, x0.coords, 1)
    xs.cmaps[xs.stride] = xs.coords
    z1 = mu.voxel_to_point(xs, z)
    g["v2p1_idx"], g["v2p1_w"] = z.idx_query[(1, 1, 1)].numpy(), z.weights[(1, 1, 1)].numpy()
    g["v2p1_vfeats"], g["v2p1_out"] = vf.numpy(), z1.F.numpy()
    x2 = F.conv3d(xs, torch.randn(8, 6, 6) * 0.3, 2, stride=2)
    z2 = mu.voxel_to_point(x2, z)
    g["s2_coords"], g["s2_feats"] = x2.coords.numpy(), x2.feats.numpy()
    g["v2p2_idx"], g["v2p2_w"] = z.idx_query[(2, 2, 2)].numpy(), z.weights[(2, 2, 2)].numpy()
    g["v2p2_out"] = z2.F.numpy()
    # point_to_voxel at stride 2
    zz = PointTensor(T(pf.copy()), z.C)
    x2v = mu.point_to_voxel(x2, zz)
    g["p2v2_idx"] = zz.additional_features["idx_query"][(2, 2, 2)].numpy()
    g["p2v2_counts"] = zz.additional_features["counts"][(2, 2, 2)].numpy()
    g["p2v2_out"] = x2v.feats.numpy()
    # raw op goldens: voxelize fwd/bwd, devoxelize fwd (the CPU devoxelize bwd twin is wrong)
    idxq = zz.additional_features["idx_query"][(2, 2, 2)]
    cnt = zz.additional_features["counts"][(2, 2, 2)]
    gv = torch.randn(x2.coords.shape[0], 4)
    g["vox_bwd_gout"] = gv.numpy()
    g["vox_bwd_gin"] = ts.backend.voxelize_backward_cpu(gv, idxq.int(), cnt, npts).numpy()
    np.savez_compressed(os.path.join(HERE, "point_voxel.npz"), **g)

    # ------------------------------------------------ spdownsample only, several shapes
    g = {}
    cc = synth_cloud(rng, 1500, 40, batch_idx=1)
    g["coords"] = cc
    for tag, (st, ks, tst) in {"s2k2": (2, 2, 1), "s2k2_t2": (2, 2, 2), "s2k3": (2, 3, 1),
                               "s221k3": ((2, 2, 1), 3, 1), "s2k3_t2": (2, 3, 2)}.items():
        src = T(cc) if tst == 1 else F.spdownsample(T(cc), 2, 2, 1)
        g[f"{tag}_in"] = src.numpy()
        g[f"{tag}_out"] = F.spdownsample(src, st, ks, tst).numpy()
    np.savez_compressed(os.path.join(HERE, "downsample.npz"), **g)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(HERE, f)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
