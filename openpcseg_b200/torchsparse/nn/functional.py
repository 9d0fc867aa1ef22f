"""torchsparse.nn.functional surface over libb2s (reference: TS/nn/functional/*.py).

Differences from the reference that are deliberate and invisible in results:
  * a kernel map is ONE fused device pass (table build + K probes per output row) that
    yields gather maps ``nbr_out`` / ``nbr_in``; the reference's ``[M, 2]`` pair list is
    derived from it in the same order and is only materialised for wgrad / inspection;
  * the convolution accumulates all kernel offsets in fp32 and writes each output row
    once (the reference accumulates across offsets in the storage dtype);
  * no host synchronisation inside conv forward/backward (the reference copies
    ``nbsizes`` to the host for every call, TS/nn/functional/conv.py:56,103).
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import backend as B
from ..tensor import SparseTensor
from ..utils import make_ntuple
from .utils import fapply, get_kernel_offsets, kernel_offsets_cached

__all__ = ["sphash", "sphashquery", "spcount", "spdownsample", "spvoxelize", "spdevoxelize",
           "calc_ti_weights", "conv3d", "relu", "leaky_relu", "KernelMap", "build_kernel_map"]


# ------------------------------------------------------------------ hash / query / count
def sphash(coords: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """TS/nn/functional/hash.py:10-37."""
    assert coords.dtype == torch.int, coords.dtype
    assert coords.ndim == 2 and coords.shape[1] == 4, coords.shape
    if offsets is None:
        return B.hash_coords(coords)
    assert offsets.dtype == torch.int, offsets.dtype
    assert offsets.ndim == 2 and offsets.shape[1] == 3, offsets.shape
    return B.kernel_hash(coords, offsets)


def sphashquery(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    """Index of each query hash in ``references`` or -1 (TS/nn/functional/query.py:8-33)."""
    shape = queries.size()
    return B.hash_query(queries.contiguous().view(-1), references.contiguous()).view(*shape)


def spcount(coords: torch.Tensor, num) -> torch.Tensor:
    """TS/nn/functional/count.py:8-16."""
    return B.count(coords.contiguous(), int(num))


def spdownsample(coords: torch.Tensor, stride=2, kernel_size=2, tensor_stride=1) -> torch.Tensor:
    """TS/nn/functional/downsample.py:11-52 (one device pass + one host sync for the count)."""
    return B.downsample_coords(coords, make_ntuple(stride, 3), make_ntuple(kernel_size, 3),
                               make_ntuple(tensor_stride, 3))


# ------------------------------------------------------------------- point <-> voxel
def _amp_half(t: torch.Tensor) -> torch.Tensor:
    """custom_fwd(cast_inputs=torch.half) of the reference: under CUDA autocast the feature
    operand runs in fp16 (TS/nn/functional/voxelize.py:13, devoxelize.py:54, conv.py:19)."""
    if torch.is_autocast_enabled() and t.is_floating_point() and t.dtype != torch.float16:
        return t.half()
    return t


class _Voxelize(Function):
    @staticmethod
    def forward(ctx, feats, idx, counts):
        ctx.aux = (idx, counts, feats.shape[0])
        return B.voxelize_forward(feats, idx, counts)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        idx, counts, n = ctx.aux
        return B.voxelize_backward(grad.contiguous(), idx, counts, n), None, None


def spvoxelize(feats: torch.Tensor, coords: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """Scatter-mean of point rows into voxel rows (TS/nn/functional/voxelize.py:10-56)."""
    return _Voxelize.apply(_amp_half(feats), coords.contiguous().int(), counts.contiguous())


class _Devoxelize(Function):
    @staticmethod
    def forward(ctx, feats, idx, weights):
        ctx.aux = (idx, weights, feats.shape[0])
        return B.devoxelize_forward(feats, idx, weights)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        idx, weights, n_vox = ctx.aux
        order = None
        if idx.shape[0] >= 2 * n_vox and grad.is_cuda:
            # many points per voxel (coarse strides): visit the points sorted by their corner-0 voxel so that the
            # kernel can merge runs in registers; the order is cached on the (per-stride cached) index tensor
            order = getattr(idx, "_b2s_order", None)
            if order is None:
                order = torch.argsort(idx[:, 0]).int()
                try:
                    idx._b2s_order = order
                except AttributeError:
                    pass
        return B.devoxelize_backward(grad.contiguous(), idx, weights, n_vox, order), None, None


def spdevoxelize(feats: torch.Tensor, coords: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Trilinear gather voxels -> points (TS/nn/functional/devoxelize.py:51-98); weights stay fp32."""
    return _Devoxelize.apply(_amp_half(feats), coords.contiguous().int(),
                             weights.detach().float().contiguous())


def calc_ti_weights(coords: torch.Tensor, idx_query: torch.Tensor, scale: float = 1) -> torch.Tensor:
    """fp32 [8, N] trilinear weights (TS/nn/functional/devoxelize.py:10-48), one kernel."""
    with torch.no_grad():
        return B.ti_weights(coords, idx_query, float(scale))


# ----------------------------------------------------------------------- kernel map
class KernelMap:
    """Device-resident rule set of one (stride, kernel, conv-stride, dilation) key.

    Behaves like the reference's ``[nbmaps, nbsizes, (N_in, N_out)]`` list
    (TS/nn/functional/conv.py:174-176) when indexed or unpacked; the int64 ``nbmaps``
    view is materialised lazily (one host sync) because the kernels here consume the
    gather maps, their step tables and the padded int32 pair buffer directly.
    """

    def __init__(self, nbr_out, nbr_in, nbsizes, sizes: Tuple[int, int], symmetric: bool,
                 mask_out=None, mask_in=None, coords=None, coord_shift: int = 0):
        self.nbr_out = nbr_out          # int32 [K, N_out]: input row per (offset, output row)
        self.nbr_in = nbr_in            # int32 [K, N_in ]: output row per (offset, input row) | None
        self.mask_out = mask_out        # active-offset bits per 128-row tile of nbr_out (API row order)
        self.mask_in = mask_in          # ... of nbr_in
        self.nbsizes32 = nbsizes        # int32 [K] on device
        self.sizes = sizes
        self.symmetric = symmetric      # nbr_in[k] == nbr_out[K-1-k] (submanifold, odd kernel)
        self.kvol = nbr_out.shape[0]
        self._coords, self._shift = coords, coord_shift
        # tile composition of the tensor-core kernels (symmetric maps): launch row j = map column perm[j],
        # rows with equal neighbourhood patterns share tiles (see _tile_order / b2s_tile_order_key)
        self._perm_ready = False
        self.tile_perm = self._row_bits = None
        self._steps = {}                # (which map, tile_rows) -> step table
        self._pairs = None              # (int32 [K*N_out, 2] padded, int64 [1] total)
        self._chunked = None            # (pairs in (row range, offset) order, segment sizes)
        self._ref = None

    def pairs(self):
        if self._pairs is None:
            self._pairs = B.kmap_pairs(self.nbr_out)
        return self._pairs

    def wgrad_pairs(self, feats: torch.Tensor):
        """(pairs, sizes) for the weight gradient: the reference-order list, or (B2S_WGRAD_CHUNKED=1, experimental)
        the list in (row range, offset) order over a spatial (batch, z, x, y) order of the rows, meant to keep one
        range's X / dY rows in L2 across its 27 offsets (the plain list re-reads rows from HBM once per offset
        when the level outgrows L2: ncu 416 MB read vs 160 MB algorithmic at batch 4, stride 1)."""
        n_out = self.nbr_out.shape[1]
        n_chunks = min(-(-n_out // _WGRAD_CHUNK_ROWS), 1024 // self.kvol)
        if n_chunks <= 1 or feats.dtype != torch.float16 or not _WGRAD_CHUNKED:
            return self.pairs()[0], self.nbsizes32
        if self._chunked is None:
            perm = _tile_order(self._coords) if (self.symmetric and self._coords is not None) else None
            pairs, seg, _ = B.kmap_pairs_chunked(self.nbr_out, perm, n_chunks)
            self._chunked = (pairs, seg)
        return self._chunked

    def total_hint(self):
        """Device scalar with the pair count M (used by the measurement hooks only)."""
        return self.pairs()[1] if B.PROFILER is not None else None

    def _ensure_perm(self):
        if self._perm_ready:
            return
        self._perm_ready = True
        n = self.nbr_out.shape[1]
        if not self.symmetric or _TILE_ORDER == "none" or n <= 256 or self._coords is None:
            return
        if _TILE_ORDER == "mask" and self.kvol <= 27:
            # rows with the same neighbourhood pattern share tiles (rarest offsets first), see
            # b2s_tile_order_key: 25 % of the (tile, offset) steps stay active at stride 1 vs 61 % for (z,x,y)
            keys, self._row_bits = B.tile_order_key(self.nbr_out, self.nbsizes32, self._coords, self._shift)
            perm = torch.argsort(keys).int()
            if _TILE_SPATIAL:
                perm = _spatial_tile_order(perm, self._coords, self._shift)
            self.tile_perm = perm
        else:
            self.tile_perm = _tile_order(self._coords)

    def gather_args(self, which: str, feats: torch.Tensor, c_red: int, c_res: int):
        """Keyword arguments of B.conv_gather_gemm for computing the rows of side ``which`` ("out": one row
        per output coordinate from input rows; "in": one row per input coordinate from output rows) plus the
        row count: step table + tile order on the fp16 tensor-core path, the plain gather map otherwise."""
        n_rows = self.sizes[1] if which == "out" else self.sizes[0]
        base, flip = ("out", True) if (which == "in" and self.symmetric) else (which, False)
        nbr = self.nbr_out if base == "out" else self.nbr_in
        if B.conv_steps_supported(feats, c_red, c_res):
            perm = None
            if base == "out":
                self._ensure_perm()
                perm = self.tile_perm
            tr = B.conv_tile_rows(c_res, n_rows)
            steps = self._steps.get((base, tr))
            if steps is None:
                steps = B.tile_steps(nbr, perm, self._row_bits if base == "out" else None, tr)
                self._steps[(base, tr)] = steps
            return n_rows, dict(nbr=None, steps=steps, row_perm=perm, flip_k=flip)
        mask = self.mask_out if base == "out" else self.mask_in
        return n_rows, dict(nbr=nbr, tile_mask=mask, row_perm=None, flip_k=flip)

    def reference_format(self):
        if self._ref is None:
            pairs, total = self.pairs()
            m = int(total.item())
            self._ref = [pairs[:m].long(), self.nbsizes32.long(), self.sizes]
        return self._ref

    def __getitem__(self, i):
        return self.reference_format()[i]

    def __iter__(self):
        return iter(self.reference_format())

    def __len__(self):
        return 3


_TILE_ORDER = os.environ.get("B2S_TILE_ORDER", "mask")      # "mask" | "zxy" | "none"
# opt-in: measured SLOWER than the plain list (profiles/r2_wgrad_chunked.txt: stride-1 level at batch 16
# 568 -> 701 us) - the plain (offset, ascending row) list reads dY sequentially, the range-major spatial order
# turns both operands into scattered gathers and doubles the number of small work units
_WGRAD_CHUNKED = os.environ.get("B2S_WGRAD_CHUNKED", "0") == "1"
_WGRAD_CHUNK_ROWS = int(os.environ.get("B2S_WGRAD_CHUNK_ROWS", 65536))
_TILE_SPATIAL = os.environ.get("B2S_TILE_SPATIAL", "0") == "1"
_TILE_SPATIAL_BITS = int(os.environ.get("B2S_TILE_SPATIAL_BITS", 5))      # block = 2^bits cells of the level


def _spatial_tile_order(perm: torch.Tensor, coords: torch.Tensor, shift: int) -> torch.Tensor:
    """Keep the COMPOSITION of the 128-row tiles (rows grouped by neighbourhood pattern: best fill) but walk the
    tiles in the order of a coarse (x, y) block of their first row instead of pattern by pattern.  The persistent
    CTAs take consecutive tiles, so the ~300 tiles in flight then cover one neighbourhood of the scene and the
    rows they gather (each needed by ~4.7 offsets / tiles) are re-used from L2; in pattern order the tiles in
    flight are scattered over the whole batch and a 16-scan level (290 MB of rows at 96 channels) streams from
    HBM once per gather."""
    n = perm.shape[0]
    g = 256                        # granule = the largest CTA tile, so 256-row tiles keep their composition too
    t = n // g
    if t < 2:
        return perm
    first = perm[: t * g: g].long()
    c = coords.index_select(0, first).long()
    blk = shift + _TILE_SPATIAL_BITS
    key = (((c[:, 0] >> blk) & 0xFFF) << 44) | (((c[:, 1] >> blk) & 0xFFF) << 32) | \
        torch.arange(t, device=perm.device, dtype=torch.int64)
    order = torch.argsort(key)
    body = perm[: t * g].view(t, g).index_select(0, order).reshape(-1)
    return torch.cat([body, perm[t * g:]])


def _tile_order(coords: torch.Tensor) -> torch.Tensor:
    """Row order in which the conv kernels walk a level: sorted by (batch, z, x, y).

    A 128-row tile then covers a compact horizontal patch of the scan, so offsets with dz != 0
    have no neighbour anywhere in most ground tiles and whole (tile, offset) steps are skipped;
    the reference row order after initial_voxelize is ascending HASH (spatially random), which
    makes every offset active in every tile.  Measured on the synthetic scan (active (tile,
    offset) fraction, profiles/): hash 1.00 -> zxy 0.61 at stride 1, 0.91 -> 0.74 at stride 4.
    Only the tile composition changes: results and the API-visible row order do not."""
    c = coords.long()
    key = (c[:, 3] << 54) | ((c[:, 2] + 131072) << 36) | ((c[:, 0] + 131072) << 18) | (c[:, 1] + 131072)
    return torch.argsort(key).int()


def build_kernel_map(in_coords: torch.Tensor, out_coords: torch.Tensor, kernel_size, in_stride,
                     dilation=1) -> KernelMap:
    kernel_size = make_ntuple(kernel_size, 3)
    offsets = kernel_offsets_cached(kernel_size, stride=in_stride, dilation=dilation, device=in_coords.device)
    same = (in_coords is out_coords) or (in_coords.data_ptr() == out_coords.data_ptr()
                                          and in_coords.shape == out_coords.shape)
    symmetric = bool(same and all(k % 2 == 1 for k in kernel_size))
    nbr_out, nbr_in, nbsizes, mask_out, mask_in = B.kmap_build(in_coords, out_coords, offsets,
                                                               want_nbr_in=not symmetric)
    s0 = int(in_stride[0]) if not isinstance(in_stride, int) else int(in_stride)
    return KernelMap(nbr_out, nbr_in, nbsizes, (in_coords.shape[0], out_coords.shape[0]), symmetric,
                     mask_out, mask_in, coords=out_coords, coord_shift=max(s0.bit_length() - 1, 0))


# ---------------------------------------------------------------------- convolution
def _weight_cache(weight: torch.Tensor, slot: str, make):
    """Derived copies of a parameter (fp16 cast, K-major transpose), cached ON the parameter object until it is
    next updated in place or re-bound (version counter + storage pointer + shape are the key)."""
    try:
        key = (weight._version, weight.data_ptr(), tuple(weight.shape), weight.device)
    except RuntimeError:                                    # inference tensors have no version counter
        return make()
    hit = weight.__dict__.get(slot) if hasattr(weight, "__dict__") else None
    if hit is not None and hit[0] == key:
        return hit[1]
    val = make()
    try:
        setattr(weight, slot, (key, val))
    except AttributeError:
        pass
    return val


class _WeightBank:
    """fp16 operand copies (parameter layout for the input gradient, K-major for the forward pass) of every fp32
    conv PARAMETER seen on a device, kept in persistent buffers and refreshed by ONE launch (b2s_weights_refresh)
    the first time a parameter is looked up after the optimizer moved its version.  Per-parameter refresh cost
    126 small launches per MinkUNet-34 step (cast + transpose per conv), each followed by a CPU-bound gap on
    the device (scripts/step_timeline.py: ~2 ms idle per step around them)."""

    class Entry:
        __slots__ = ("ref", "cast", "kmajor", "key", "k", "c_in", "c_out")

    def __init__(self, device):
        self.device = device
        self.entries = []
        self.by_id = {}                 # id(parameter) -> entry (nothing is attached to the parameter itself: its
        #                                 __dict__ is pickled by torch.save(model))
        self.table = None               # (device int64 [n, 5], total units, the entries it describes)

    @staticmethod
    def _units(e):
        return e.k * (-(-e.c_in // 32)) * (-(-e.c_out // 32))

    def _launch(self, entries, table=None):
        if table is None:
            import numpy as np
            rows, start = [], 0
            for e in entries:
                w = e.ref()
                rows.append((w.data_ptr(), e.cast.data_ptr(), e.kmajor.data_ptr(), e.k | (e.c_in << 32),
                             e.c_out | (start << 32)))
                start += self._units(e)
            table = (torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(self.device), start)
        B.weights_refresh(table[0], table[1])
        for e in entries:
            w = e.ref()
            e.key = (w._version, w.data_ptr())
        return table

    def lookup(self, weight):
        e = self.by_id.get(id(weight))
        if e is None or e.ref() is not weight:              # (ids are re-used once a parameter is gone)
            e = _WeightBank.Entry()
            e.ref = weakref.ref(weight)
            shape3 = tuple(weight.shape) if weight.ndim == 3 else (1,) + tuple(weight.shape)
            e.k, e.c_in, e.c_out = shape3
            e.cast = torch.empty(weight.shape, dtype=torch.float16, device=weight.device)
            e.kmajor = torch.empty((e.k, e.c_out, e.c_in), dtype=torch.float16, device=weight.device)
            e.key = None
            self.by_id[id(weight)] = e
            self.entries.append(e)
            self.table = None
            self._launch([e])                               # first sight: this parameter alone
        elif e.key != (weight._version, weight.data_ptr()):
            self.refresh_all()
        return e

    def refresh_all(self):
        live = [e for e in self.entries if e.ref() is not None]
        moved = any(e.key is None or e.key[1] != e.ref().data_ptr() for e in live)
        if len(live) != len(self.entries) or moved or self.table is None or self.table[2] != len(live):
            self.entries = live
            self.by_id = {id(e.ref()): e for e in live}
            self.table = None
        t = self._launch(live, None if self.table is None else self.table[:2])
        self.table = (t[0], t[1], len(live))


_WEIGHT_BANKS = {}
_WEIGHT_BANK = os.environ.get("B2S_WEIGHT_BANK", "1") != "0"


def _bankable(weight: torch.Tensor, dtype: torch.dtype) -> bool:
    return (_WEIGHT_BANK and dtype == torch.float16 and weight.dtype == torch.float32 and weight.is_cuda
            and isinstance(weight, torch.nn.Parameter) and weight.ndim in (2, 3) and weight.is_contiguous())


def _banked(weight: torch.Tensor, dtype: torch.dtype):
    """The bank entry of ``weight`` when it is an fp32 CUDA conv parameter used with fp16 features, else None."""
    if not _bankable(weight, dtype):
        return None
    bank = _WEIGHT_BANKS.get(weight.device)
    if bank is None:
        bank = _WEIGHT_BANKS[weight.device] = _WeightBank(weight.device)
    return bank.lookup(weight)


def _cast_weight(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp32 master weight -> feature dtype (the reference re-casts on every call through custom_fwd,
    TS/nn/functional/conv.py:19)."""
    if weight.dtype == dtype:
        return weight
    e = _banked(weight, dtype)
    if e is not None:
        return e.cast
    return _weight_cache(weight, "_b2s_cast_" + str(dtype).split(".")[-1], lambda: weight.detach().to(dtype))


def _kmajor_weight(weight: torch.Tensor, w_cast: torch.Tensor):
    """The forward pass's K-major operand [K, C_out, C_in] of an fp16 weight, cached with the parameter (the
    first revision transposed the weight inside every forward call: 63 extra launches per step)."""
    if w_cast.dtype != torch.float16:
        return None
    if _bankable(weight, torch.float16):
        bank = _WEIGHT_BANKS.get(weight.device)
        e = bank.by_id.get(id(weight)) if bank is not None else None
        if e is not None and e.cast is w_cast and e.ref() is weight:
            return e.kmajor
    w3 = w_cast if w_cast.ndim == 3 else w_cast.unsqueeze(0)
    return _weight_cache(weight, "_b2s_kmajor", lambda: B.weight_to_kmajor(w3.contiguous()))


def _conv_rows(feats, w, kmap: Optional[KernelMap], which: str, transpose_w: bool, w_kmajor=None, hint=None,
               bn_sums=None):
    """Rows of side ``which`` of ``kmap`` (None: identity map / 1x1 conv) = gather-GEMM of ``feats`` with ``w``
    [K, C_in, C_out]; result widths above the kernel's 512 TMEM columns are split into column blocks of the
    weight (e.g. the input gradient of RPVNet's 672 -> 448 decoder conv)."""
    c_red, c_res = (w.shape[2], w.shape[1]) if transpose_w else (w.shape[1], w.shape[2])

    def run(wb, width, kmajor, sums):
        if kmap is None:
            return B.conv_gather_gemm(feats, wb, None, feats.shape[0], transpose_w, False, pairs_hint=hint,
                                      weight_kmajor=kmajor, bn_sums=sums)
        n_rows, kw = kmap.gather_args(which, feats, c_red, width)
        return B.conv_gather_gemm(feats, wb, n_rows=n_rows, transpose_w=transpose_w, pairs_hint=hint,
                                  weight_kmajor=kmajor, bn_sums=sums if "steps" in kw else None, **kw)

    if c_res <= 512 or feats.dtype != torch.float16:
        return run(w, c_res, w_kmajor, bn_sums)
    assert bn_sums is None
    n_blk = -(-c_res // 512)
    step = -(-c_res // (n_blk * 32)) * 32
    outs = []
    for c0 in range(0, c_res, step):
        wb = (w[:, c0:c0 + step, :] if transpose_w else w[:, :, c0:c0 + step]).contiguous()
        outs.append(run(wb, wb.shape[1] if transpose_w else wb.shape[2], None, None))
    return torch.cat(outs, dim=1)


class ConvolutionFunction(Function):
    """out = sum_k gather(in, map_k) @ W[k]  (TS/nn/functional/conv.py:16-119).

    ``weight`` keeps its parameter dtype (fp32 master); it is cast to the feature dtype
    here and its gradient is returned in fp32 straight from the wgrad kernel.
    """

    @staticmethod
    def forward(ctx, feats, weight, kmap: KernelMap, transposed: bool, bn_sums=None):
        feats = feats.contiguous()
        w = _cast_weight(weight, feats.dtype)
        out = _conv_rows(feats, w, kmap, "in" if transposed else "out", False, _kmajor_weight(weight, w),
                         kmap.total_hint(), bn_sums)
        ctx.save_for_backward(feats, weight)
        ctx.kmap, ctx.transposed = kmap, transposed
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        feats, weight = ctx.saved_tensors
        kmap, transposed = ctx.kmap, ctx.transposed
        grad_out = grad_out.contiguous()
        w = _cast_weight(weight, feats.dtype)
        grad_in = grad_w = None
        hint = kmap.total_hint()
        if ctx.needs_input_grad[0]:
            grad_in = _conv_rows(grad_out, w, kmap, "out" if transposed else "in", True, None, hint)
        if ctx.needs_input_grad[1]:
            pairs, sizes = kmap.wgrad_pairs(feats)
            grad_w = B.conv_wgrad(feats, grad_out, kmap.kvol, pairs, sizes, transposed, pairs_hint=hint)
            grad_w = grad_w.to(weight.dtype)
        return grad_in, grad_w, None, None, None


class _DenseConv(Function):
    """1x1x1 conv = per-row GEMM through the same kernel family (identity map)."""

    @staticmethod
    def forward(ctx, feats, weight):
        feats = feats.contiguous()
        ctx.save_for_backward(feats, weight)
        w = _cast_weight(weight, feats.dtype)
        return _conv_rows(feats, w if w.ndim == 3 else w.unsqueeze(0), None, "out", False, _kmajor_weight(weight, w))

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        feats, weight = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_in = grad_w = None
        if ctx.needs_input_grad[0]:
            w = _cast_weight(weight, feats.dtype)
            grad_in = _conv_rows(grad_out, w if w.ndim == 3 else w.unsqueeze(0), None, "in", True)
        if ctx.needs_input_grad[1]:
            grad_w = B.conv_wgrad(feats, grad_out, 1, None, None, False)[0].to(weight.dtype)
        return grad_in, grad_w


def _pad_for_tensor_cores(feats: torch.Tensor, weight: torch.Tensor):
    """fp16 only: zero-pad C_in and C_out to multiples of 32 so that narrow or odd layers (the 4-channel
    stem, 20-class heads, the 56/112/168-channel layers of the cr 1.75 models) run on the tcgen05 kernels
    in all three passes (forward reduces over C_in, the input gradient over C_out) instead of the
    CUDA-core family; zero channels change nothing and autograd slices the gradients back."""
    if feats.dtype != torch.float16:
        return feats, weight, None
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    pad_in, pad_out = (-c_in) % 32, (-c_out) % 32
    if pad_in == 0 and pad_out == 0:
        return feats, weight, None
    feats = torch.nn.functional.pad(feats, (0, pad_in))
    weight = torch.nn.functional.pad(weight, (0, pad_out, 0, pad_in))
    return feats, weight, (c_out if pad_out else None)


_AUTO_BN_SUMS = os.environ.get("B2S_AUTO_BN_SUMS", "1") != "0"
_ZERO_CHUNK = 1 << 16                   # doubles per memset: ~2 steps of MinkUNet-34's 55 statistics buffers
_zero_arena = {}                        # device -> [chunk, next free element]


def zero_sums(c: int, device) -> torch.Tensor:
    """fp64 zeros [2, c] for a conv epilogue's batch-norm statistics, cut from a chunk zeroed by ONE memset
    (a ``torch.zeros`` per conv was 55 fill launches per MinkUNet-34 step).  Every slice is handed out once."""
    n = 2 * c
    device = torch.device(device)
    slot = _zero_arena.get(device)
    if slot is None or slot[1] + n > slot[0].numel():
        slot = _zero_arena[device] = [torch.zeros(max(_ZERO_CHUNK, n), dtype=torch.float64, device=device), 0]
    out = slot[0][slot[1]: slot[1] + n].view(2, c)
    slot[1] += n
    return out


def _auto_sums(bn_sums, feats, weight, keep_out, bias):
    """The fp64 [2, C_out] statistics buffer the conv epilogue should fill, or None.  A caller may pass one
    (``conv3d(bn_sums=...)``); otherwise every bias-free training-time conv on the step-table kernel gets one
    speculatively - the reference's blocks are conv -> BatchNorm (minkunet.py:66-72, 100-114) and the batch norm
    that follows then skips its own pass over [N, C]; a conv without a norm behind it wastes one memset."""
    ok = (keep_out is None and bias is None and weight.ndim == 3 and weight.shape[2] <= 512
          and B.conv_steps_supported(feats, weight.shape[1], weight.shape[2]))
    if not ok:
        return None
    if bn_sums is None and _AUTO_BN_SUMS and torch.is_grad_enabled() and weight.requires_grad:
        bn_sums = zero_sums(weight.shape[2], feats.device)
    return bn_sums


def _fused_sums(x: torch.Tensor):
    """Statistics attached to ``x`` by the conv that produced it, if ``x`` was not modified since."""
    tag = getattr(x, "_b2s_sums", None)
    if tag is None or tag[1] != x._version or tag[0].shape[1] != x.shape[1]:
        return None
    return tag[0]


def conv3d(input: SparseTensor, weight: torch.Tensor,
           kernel_size: Union[int, List[int], Tuple[int, ...]], bias: Optional[torch.Tensor] = None,
           stride: Union[int, List[int], Tuple[int, ...]] = 1,
           dilation: Union[int, Tuple[int, ...]] = 1, transposed: bool = False,
           bn_sums: Optional[torch.Tensor] = None) -> SparseTensor:
    """Sparse 3-D convolution with the reference's coordinate / map caching rules
    (TS/nn/functional/conv.py:122-205).

    ``bn_sums`` (extension, fp64 zeros [2, C_out]): when the fp16 step-table kernel runs this conv, its epilogue
    adds the per-channel sum / sum of squares of the result rows into it and the result's ``feats`` carries it as
    ``_b2s_sums`` - ``batch_norm_act`` then skips its statistics pass."""
    kernel_size = make_ntuple(kernel_size, ndim=3)
    stride = make_ntuple(stride, ndim=3)
    dilation = make_ntuple(dilation, ndim=3)
    feats, weight, keep_out = _pad_for_tensor_cores(_amp_half(input.feats), weight)
    ones = (1, 1, 1)

    if kernel_size == ones and stride == ones and dilation == ones:
        out_stride, out_coords = input.stride, input.coords
        out_feats = _DenseConv.apply(feats, weight)
    elif not transposed:
        out_stride = tuple(input.stride[a] * stride[a] for a in range(3))
        if out_stride in input.cmaps:
            out_coords = input.cmaps[out_stride]
        elif stride == ones:
            out_coords = input.coords
        else:
            out_coords = spdownsample(input.coords, stride, kernel_size, input.stride)
        key = (input.stride, kernel_size, stride, dilation)
        if key not in input.kmaps:
            input.kmaps[key] = build_kernel_map(input.coords, out_coords, kernel_size, input.stride,
                                                dilation)
        bn_sums = _auto_sums(bn_sums, feats, weight, keep_out, bias)
        out_feats = ConvolutionFunction.apply(feats, weight, input.kmaps[key], False, bn_sums)
        if bn_sums is not None:
            out_feats._b2s_sums = (bn_sums, out_feats._version)
    else:
        out_stride = tuple(input.stride[a] // stride[a] for a in range(3))
        out_coords = input.cmaps[out_stride]
        kmap = input.kmaps[(out_stride, kernel_size, stride, dilation)]
        bn_sums = _auto_sums(bn_sums, feats, weight, keep_out, bias)
        out_feats = ConvolutionFunction.apply(feats, weight, kmap, True, bn_sums)
        if bn_sums is not None:
            out_feats._b2s_sums = (bn_sums, out_feats._version)

    if keep_out is not None:
        # contiguous: a strided [N, C] view sends torch's batch norm to its generic (non channels-last)
        # kernels - 14 ms per backward launch on the cr 1.75 models
        out_feats = out_feats[:, :keep_out].contiguous()
    if bias is not None:
        out_feats = out_feats + bias.to(out_feats.dtype)

    out = SparseTensor(out_feats, out_coords, out_stride)
    out.cmaps = input.cmaps
    out.cmaps.setdefault(out_stride, out_coords)
    out.kmaps = input.kmaps
    return out


# ----------------------------------------------------------------------- activations
def relu(input: SparseTensor, inplace: bool = True) -> SparseTensor:
    return fapply(input, torch.nn.functional.relu, inplace=inplace)


def leaky_relu(input: SparseTensor, negative_slope: float = 0.1, inplace: bool = True) -> SparseTensor:
    return fapply(input, torch.nn.functional.leaky_relu, negative_slope=negative_slope,
                  inplace=inplace)


# ------------------------------------------------------ fused batch norm (+add, +ReLU)
class _BatchNormAct(Function):
    """y = act(batch_norm_train(x) [+ residual]) in two kernels forward / two backward."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, eps, momentum, relu, sums=None):
        y, mean, invstd, scale_shift = B.bn_forward(x, residual, gamma, beta, running_mean, running_var, eps,
                                                    momentum, relu, sums)
        has_res = residual is not None
        # ReLU mask of the backward pass: from y when a residual was added, else recomputed from x (saves a
        # read of y in both backward kernels and does not keep y alive for autograd)
        ctx.save_for_backward(x, y if (relu and has_res) else None, mean, invstd, gamma,
                              scale_shift if (relu and not has_res) else None)
        ctx.relu, ctx.has_res = relu, has_res
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, y, mean, invstd, gamma, scale_shift = ctx.saved_tensors
        dx, dres, dgamma, dbeta = B.bn_backward(dy, y, x, mean, invstd, gamma, ctx.relu,
                                                ctx.has_res and ctx.needs_input_grad[1], scale_shift)
        return (dx, dres, dgamma.to(gamma.dtype) if gamma is not None else None,
                dbeta.to(gamma.dtype) if gamma is not None else None, None, None, None, None, None, None)


class _SyncBatchNormAct(Function):
    """Synchronised (cross-rank) variant of _BatchNormAct: the same kernels with one all-reduce of the fp64
    [2C + 1] statistics (sums + row count) in forward and one of the [2C] gradient sums in backward - instead of
    torch.nn.SyncBatchNorm's unfused all_gather / all_reduce sequence per layer (reference: IF_DIST True,
    minkunet.py:23-25; torch/nn/modules/_functions.py SyncBatchNorm).  Weight / bias gradients are the LOCAL
    sums, as in torch: DistributedDataParallel averages them with the other parameters."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, eps, momentum, relu, group, sums=None):
        import torch.distributed as dist
        c = x.shape[1]
        buf = B.bn_stats(x, extra=1) if sums is None else torch.cat([sums.reshape(-1), sums.new_zeros(1)])
        buf[2 * c] = float(x.shape[0])
        dist.all_reduce(buf, group=group)
        y, mean, invstd, scale_shift = B.bn_forward_global(x, residual, gamma, beta, running_mean, running_var,
                                                           eps, momentum, relu, buf)
        has_res = residual is not None
        ctx.save_for_backward(x, y if (relu and has_res) else None, mean, invstd, gamma, buf[2 * c:],
                              scale_shift if (relu and not has_res) else None)
        ctx.relu, ctx.has_res, ctx.group = relu, has_res, group
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        import torch.distributed as dist
        x, y, mean, invstd, gamma, n_total, scale_shift = ctx.saved_tensors
        local = B.bn_backward_reduce(dy, y, x, mean, invstd, ctx.relu, scale_shift)
        glob = local.clone()
        dist.all_reduce(glob, group=ctx.group)
        dx, dres = B.bn_backward_apply(dy, y, x, mean, invstd, gamma, ctx.relu,
                                       ctx.has_res and ctx.needs_input_grad[1], glob, n_total, scale_shift)
        dgamma = local[1].float().to(gamma.dtype) if gamma is not None else None
        dbeta = local[0].float().to(gamma.dtype) if gamma is not None else None
        return (dx, dres, dgamma, dbeta) + (None,) * 7


def batch_norm_act(x: torch.Tensor, bn: torch.nn.modules.batchnorm._BatchNorm, relu: bool = False,
                   residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused training-mode batch norm of [N, C] rows with optional residual add and ReLU, same
    numerics contract as ``relu(bn(x) + residual)``.  Falls back to the stock modules in eval mode,
    for SyncBatchNorm under a process group, or for channel counts the kernels do not tile."""
    sync = isinstance(bn, torch.nn.SyncBatchNorm) and torch.distributed.is_available() and \
        torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    usable = bn.training and bn.track_running_stats and bn.momentum is not None and B.bn_supported(x) \
        and (residual is None or residual.dtype == x.dtype)
    if usable and sync:
        with torch.no_grad():
            bn.num_batches_tracked += 1
        return _SyncBatchNormAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                                       bn.momentum, relu, bn.process_group, _fused_sums(x))
    if not usable:
        # the dense implementation of the module's class on the [N, C] rows (the sparse wrappers'
        # own forward expects a SparseTensor)
        dense = torch.nn.SyncBatchNorm if isinstance(bn, torch.nn.SyncBatchNorm) else torch.nn.BatchNorm1d
        y = dense.forward(bn, x)
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y
    with torch.no_grad():
        bn.num_batches_tracked += 1
    return _BatchNormAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                               bn.momentum, relu, _fused_sums(x))


def batch_norm_fusable(x_dtype: torch.dtype, bn: torch.nn.modules.batchnorm._BatchNorm) -> bool:
    """Whether ``batch_norm_act`` will take its fused training path for this module (then the producing conv may
    accumulate the statistics, ``conv3d(bn_sums=...)``)."""
    return bool(bn.training and bn.track_running_stats and bn.momentum is not None and x_dtype == torch.float16)
