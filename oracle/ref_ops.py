"""CPU oracle for the sparse-voxel hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy / CPU torch, the algorithms of the
reference backend (torchsparse 1.4.0 as bundled in
``/root/reference/package/torchsparse.zip``; paths below written ``TS/...``
are relative to ``torchsparse/torchsparse/`` inside that zip).  It exists so the
CUDA path can be checked for parity.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it; the product package ``openpcseg_b200`` never
does (and fails loudly without its CUDA library).

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``tests/golden/*.npz`` were produced by
``tests/golden/make_golden.py`` importing the reference's compiled CPU backend
in the build container, and ``tests/test_oracle_golden.py`` checks every
function here against them.  Known bugs of the reference *CPU twin* (not of its
CUDA path) that the oracle deliberately does not reproduce are listed in
DESIGN.md ("oracle caveats").
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

_FNV_OFFSET = np.uint64(14695981039346656037)
_FNV_PRIME = np.uint64(1099511628211)
_LOW60 = np.uint64(0x0FFFFFFFFFFFFFFF)

IntOrTriple = Union[int, Sequence[int]]


def make_ntuple(x: IntOrTriple, ndim: int = 3) -> Tuple[int, ...]:
    """TS/utils/utils.py:9-20."""
    if isinstance(x, (int, np.integer)):
        return tuple(int(x) for _ in range(ndim))
    x = tuple(int(v) for v in x)
    assert len(x) == ndim, x
    return x


# --------------------------------------------------------------------------- hash
def sphash(coords: np.ndarray, offsets: Optional[np.ndarray] = None) -> np.ndarray:
    """60-bit folded FNV-1a-per-word hash of (x, y, z, batch).

    Follows TS/backend/hash/hash_cuda.cu:10-23 (plain) and :27-56 (with kernel
    offsets, result laid out [K, N]).  Uses the CUDA semantics for the batch
    word (each point's own batch index), not the CPU twin's ``data[3]`` slip
    (TS/backend/hash/hash_cpu.cpp:29).
    """
    c = np.ascontiguousarray(coords, dtype=np.int32)
    assert c.ndim == 2 and c.shape[1] == 4, c.shape
    if offsets is None:
        words = c[None, :, :]                               # [1, N, 4]
    else:
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        assert off.ndim == 2 and off.shape[1] == 3, off.shape
        shifted = c[None, :, :3] + off[:, None, :]          # int32 wrap-around add
        batch = np.broadcast_to(c[None, :, 3:], (off.shape[0], c.shape[0], 1))
        words = np.concatenate([shifted, batch], axis=2)    # [K, N, 4]
    u = words.astype(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = np.full(u.shape[:2], _FNV_OFFSET, dtype=np.uint64)
        for j in range(4):
            h = (h ^ u[:, :, j]) * _FNV_PRIME
    h = (h >> np.uint64(60)) ^ (h & _LOW60)
    h = h.astype(np.int64)
    return h[0] if offsets is None else h


def get_kernel_offsets(size: IntOrTriple, stride: IntOrTriple = 1,
                       dilation: IntOrTriple = 1) -> np.ndarray:
    """Kernel offsets int32 [K, 3]; TS/nn/utils/kernel.py:11-32.

    Odd kernel volume: x varies fastest (z outermost); even volume: z fastest.
    """
    size, stride, dilation = make_ntuple(size), make_ntuple(stride), make_ntuple(dilation)
    axes = [np.arange((-size[a]) // 2 + 1, size[a] // 2 + 1) * stride[a] * dilation[a]
            for a in range(3)]
    if int(np.prod(size)) % 2 == 1:
        zz, yy, xx = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    else:
        xx, yy, zz = np.meshgrid(axes[0], axes[1], axes[2], indexing="ij")
    return np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1).astype(np.int32)


def sphashquery(queries: np.ndarray, references: np.ndarray) -> np.ndarray:
    """Index of each query key in ``references`` or -1.

    TS/nn/functional/query.py:8-33 + TS/backend/others/query_cpu.cpp:12-37
    (value stored = row index + 1, miss = 0, python subtracts 1).  Duplicate
    reference keys: first insertion wins (query_cpu.cpp:22-26).
    """
    q = np.asarray(queries, dtype=np.int64)
    r = np.asarray(references, dtype=np.int64).ravel()
    out = np.full(q.size, -1, dtype=np.int64)
    if r.size:
        order = np.argsort(r, kind="stable")              # stable => first occurrence first
        rs = r[order]
        pos = np.searchsorted(rs, q.ravel(), side="left")
        pos_c = np.minimum(pos, r.size - 1)
        hit = rs[pos_c] == q.ravel()
        out[hit] = order[pos_c[hit]]
    return out.reshape(q.shape)


def spcount(idx: np.ndarray, num: int) -> np.ndarray:
    """int32 histogram of the non-negative entries; TS/backend/others/count_cuda.cu:10-16."""
    idx = np.asarray(idx).ravel()
    return np.bincount(idx[idx >= 0], minlength=int(num)).astype(np.int32)[: int(num)]


# --------------------------------------------------------------- coordinates / maps
def spdownsample(coords: np.ndarray, stride: IntOrTriple = 2, kernel_size: IntOrTriple = 2,
                 tensor_stride: IntOrTriple = 1) -> np.ndarray:
    """Output coordinates of a strided conv; TS/nn/functional/downsample.py:11-52.

    Fast path (every stride equals 1 or the kernel size): snap to the coarser
    grid with truncation toward zero.  Slow path: expand by the kernel offsets,
    keep candidates on the coarse grid and >= the per-axis minimum.  Result is
    the sorted set of unique rows, ordered by (batch, x, y, z).
    """
    stride, kernel_size, tensor_stride = (make_ntuple(stride), make_ntuple(kernel_size),
                                          make_ntuple(tensor_stride))
    c = np.ascontiguousarray(coords, dtype=np.int32)
    step = np.array([stride[a] * tensor_stride[a] for a in range(3)], dtype=np.int32)
    if all(stride[a] in (1, kernel_size[a]) for a in range(3)):
        out = c.copy()
        q = np.trunc(c[:, :3].astype(np.float32) / step.astype(np.float32))
        out[:, :3] = (q * step.astype(np.float32)).astype(np.int32)
    else:
        off = get_kernel_offsets(kernel_size, tensor_stride)
        cmin = c[:, :3].min(axis=0, keepdims=True)
        xyz = (c[:, None, :3] + off[None, :, :]).reshape(-1, 3)
        b = np.repeat(c[:, 3], off.shape[0])
        keep = np.all((np.mod(xyz, step) == 0) & (xyz >= cmin), axis=1)
        out = np.concatenate([xyz[keep], b[keep, None]], axis=1).astype(np.int32)
    bxyz = np.unique(out[:, [3, 0, 1, 2]], axis=0)
    return np.ascontiguousarray(bxyz[:, [1, 2, 3, 0]])


def build_kmap(in_coords: np.ndarray, out_coords: np.ndarray, kernel_size: IntOrTriple,
               in_stride: IntOrTriple = 1, dilation: IntOrTriple = 1
               ) -> Tuple[np.ndarray, np.ndarray]:
    """(nbmaps int64 [M, 2] = (in_idx, out_idx), nbsizes int64 [K]).

    TS/nn/functional/conv.py:156-176: pairs satisfy in_coord = out_coord +
    offset[k]; grouped by k ascending, out_idx ascending inside a group.
    """
    off = get_kernel_offsets(kernel_size, in_stride, dilation)
    refs = sphash(in_coords)
    res = sphashquery(sphash(out_coords, off), refs)        # [K, N_out]
    hit = res != -1
    nbsizes = hit.sum(axis=1).astype(np.int64)
    k_idx, o_idx = np.nonzero(hit)
    nbmaps = np.stack([res[k_idx, o_idx], o_idx], axis=1).astype(np.int64)
    return nbmaps.reshape(-1, 2), nbsizes


# -------------------------------------------------------------------- convolution
def conv_forward(feats: np.ndarray, weight: np.ndarray, nbmaps: np.ndarray, nbsizes: np.ndarray,
                 sizes: Tuple[int, int], transposed: bool = False) -> np.ndarray:
    """out[o] += in[i] @ W[k] over every pair; TS/nn/functional/conv.py:68-79
    (the pure-torch branch is the spec) and TS/backend/convolution/convolution_cpu.cpp:38-106.
    Accumulates in float64 and returns float32.
    """
    n_out = sizes[0] if transposed else sizes[1]
    out = np.zeros((n_out, weight.shape[-1]), dtype=np.float64)
    x = feats.astype(np.float64)
    w = weight.astype(np.float64)
    a = 0
    for k in range(w.shape[0]):
        b = a + int(nbsizes[k])
        src = nbmaps[a:b, 1 if transposed else 0]
        dst = nbmaps[a:b, 0 if transposed else 1]
        if b > a:
            np.add.at(out, dst, x[src] @ w[k])
        a = b
    return out.astype(np.float32)


def conv_backward(feats: np.ndarray, weight: np.ndarray, grad_out: np.ndarray, nbmaps: np.ndarray,
                  nbsizes: np.ndarray, transposed: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """(grad_feats, grad_weight); TS/backend/convolution/convolution_cuda.cu:167-278:
    per offset dX[i] += dY[o] @ W[k]^T and dW[k] = X[i]^T @ dY[o].
    """
    x, w, g = feats.astype(np.float64), weight.astype(np.float64), grad_out.astype(np.float64)
    gx = np.zeros_like(x)
    gw = np.zeros_like(w)
    a = 0
    for k in range(w.shape[0]):
        b = a + int(nbsizes[k])
        src = nbmaps[a:b, 1 if transposed else 0]
        dst = nbmaps[a:b, 0 if transposed else 1]
        if b > a:
            np.add.at(gx, src, g[dst] @ w[k].T)
            gw[k] = x[src].T @ g[dst]
        a = b
    return gx.astype(np.float32), gw.astype(np.float32)


# ------------------------------------------------------------- point <-> voxel ops
def spvoxelize_forward(feats: np.ndarray, idx: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Scatter-mean of point rows into voxel rows; TS/backend/voxelize/voxelize_cuda.cu:12-25."""
    n_vox = counts.shape[0]
    out = np.zeros((n_vox, feats.shape[1]), dtype=np.float64)
    idx = np.asarray(idx).astype(np.int64)
    ok = (idx >= 0) & (idx < n_vox)
    ok[ok] &= counts[idx[ok]] != 0
    np.add.at(out, idx[ok], feats[ok].astype(np.float64) / counts[idx[ok]][:, None])
    return out.astype(np.float32)


def spvoxelize_backward(grad_vox: np.ndarray, idx: np.ndarray, counts: np.ndarray,
                        n_pts: int) -> np.ndarray:
    """grad_pts[i] = grad_vox[idx[i]] / count; voxelize_cuda.cu:28-42."""
    n_vox = counts.shape[0]
    out = np.zeros((n_pts, grad_vox.shape[1]), dtype=np.float32)
    idx = np.asarray(idx).astype(np.int64)
    ok = (idx >= 0) & (idx < n_vox)
    ok[ok] &= counts[idx[ok]] != 0
    out[ok] = grad_vox[idx[ok]] / counts[idx[ok]][:, None].astype(np.float32)
    return out


def calc_ti_weights(coords: np.ndarray, idx_query: np.ndarray, scale: float = 1) -> np.ndarray:
    """Trilinear weights fp32 [8, N]; TS/nn/functional/devoxelize.py:10-48.

    Corner order: (x, y, z) bits with z least significant; divide by scale^3,
    zero the missing corners, renormalise by (sum + 1e-8).  All arithmetic in
    float32 like the reference.
    """
    p = coords[:, :3].astype(np.float32)
    s = np.float32(scale)
    pf = np.floor(p / s) * s if scale != 1 else np.floor(p)
    pc = pf + s
    lo = p - pf                     # distance to the floor corner
    hi = pc - p                     # distance to the ceil corner
    w = np.empty((8, p.shape[0]), dtype=np.float32)
    for corner in range(8):
        fx = lo[:, 0] if corner & 4 else hi[:, 0]
        fy = lo[:, 1] if corner & 2 else hi[:, 1]
        fz = lo[:, 2] if corner & 1 else hi[:, 2]
        w[corner] = fx * fy * fz
    if scale != 1:
        w /= np.float32(scale ** 3)
    w[idx_query == -1] = 0
    w /= w.sum(axis=0, dtype=np.float32) + np.float32(1e-8)
    return w


def spdevoxelize_forward(feats: np.ndarray, idx: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """out[p] = sum_k w[p, k] * feats[idx[p, k]] (idx < 0 contributes 0);
    TS/backend/devoxelize/devoxelize_cuda.cu:11-34."""
    idx = np.asarray(idx).astype(np.int64)
    f = feats.astype(np.float64)
    out = np.zeros((idx.shape[0], feats.shape[1]), dtype=np.float64)
    for k in range(idx.shape[1]):
        ok = idx[:, k] >= 0
        out[ok] += weights[ok, k:k + 1].astype(np.float64) * f[idx[ok, k]]
    return out.astype(np.float32)


def spdevoxelize_backward(grad_pts: np.ndarray, idx: np.ndarray, weights: np.ndarray,
                          n_vox: int) -> np.ndarray:
    """grad_vox[idx[p, k]] += w[p, k] * grad_pts[p]; devoxelize_cuda.cu:37-58.
    (The reference CPU twin of this op is wrong, devoxelize_cpu.cpp:48-53; this
    follows the CUDA kernel.)"""
    idx = np.asarray(idx).astype(np.int64)
    out = np.zeros((n_vox, grad_pts.shape[1]), dtype=np.float64)
    g = grad_pts.astype(np.float64)
    for k in range(idx.shape[1]):
        ok = idx[:, k] >= 0
        np.add.at(out, idx[ok, k], weights[ok, k:k + 1].astype(np.float64) * g[ok])
    return out.astype(np.float32)


# ------------------------------------------------- model-side helpers (callers of the path)
def initial_voxelize(pt_coords: np.ndarray, pt_feats: np.ndarray, init_res: float,
                     after_res: float):
    """Points -> stride-1 voxels, voxel order = ascending hash.

    pcseg/model/segmentor/voxel/minkunet/utils.py:11-36.  Returns
    (voxel_coords int32 [V, 4], voxel_feats fp32 [V, C], idx_query int64 [N],
    counts int32 [V], new_float_coord fp32 [N, 4]).
    """
    c = pt_coords.astype(np.float32)
    nfc = np.concatenate([(c[:, :3] * np.float32(init_res)) / np.float32(after_res), c[:, 3:4]],
                         axis=1).astype(np.float32)
    fl = np.floor(nfc)
    pc_hash = sphash(fl.astype(np.int32))
    sparse_hash = np.unique(pc_hash)
    idx_query = sphashquery(pc_hash, sparse_hash)
    counts = spcount(idx_query.astype(np.int32), sparse_hash.shape[0])
    vox_c = np.round(spvoxelize_forward(fl, idx_query, counts)).astype(np.int32)
    vox_f = spvoxelize_forward(pt_feats, idx_query, counts)
    return vox_c, vox_f, idx_query, counts, nfc


def trilinear_map(pt_coords: np.ndarray, vox_coords: np.ndarray, stride: int):
    """(idx_query int64 [N, 8], weights fp32 [N, 8]) of voxel_to_point;
    pcseg/model/segmentor/voxel/minkunet/utils.py:69-105 (nearest=False)."""
    off = get_kernel_offsets(2, stride, 1)
    base = np.concatenate([
        (np.floor(pt_coords[:, :3] / np.float32(stride)).astype(np.int32) * stride),
        pt_coords[:, 3:4].astype(np.int32)], axis=1).astype(np.int32)
    idx = sphashquery(sphash(base, off), sphash(vox_coords))
    w = calc_ti_weights(pt_coords, idx, scale=stride)
    return np.ascontiguousarray(idx.T), np.ascontiguousarray(w.T)


def point_to_voxel_map(pt_coords: np.ndarray, vox_coords: np.ndarray, stride: int):
    """(idx_query int64 [N], counts int32 [V]) of point_to_voxel; utils.py:41-64."""
    base = np.concatenate([
        (np.floor(pt_coords[:, :3] / np.float32(stride)).astype(np.int32) * stride),
        pt_coords[:, 3:4].astype(np.int32)], axis=1).astype(np.int32)
    idx = sphashquery(sphash(base), sphash(vox_coords))
    return idx, spcount(idx.astype(np.int32), vox_coords.shape[0])
