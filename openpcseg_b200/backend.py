"""Tensor-level wrappers over the C ABI (include/b2s.h).

PyTorch is plumbing here: it owns device memory (caching allocator) and the
current stream; every function unwraps ``data_ptr()`` + the stream handle and
calls libb2s.  CUDA tensors only - CPU tensors raise, there is no CPU path.

The second half of the module mirrors the names of the reference's pybind module
``torchsparse.backend`` (TS/backend/pybind_cuda.cpp:18-39) so code written
against that boundary keeps working.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check as _check

F32, F16 = 0, 1

# ---- measurement hooks (bench.py): count of libb2s kernel launches and optional per-call
# CUDA-event timing of the convolution kernels.  Off the hot path unless PROFILER is set.
STATS = {"launches": 0}
PROFILER = None          # object with .event() -> cuda Event and .record(kind, meta, start, end)


class _Timed:
    """Brackets one kernel launch with CUDA events on the current stream when profiling."""

    __slots__ = ("kind", "meta", "start")

    def __init__(self, kind: str, meta: dict):
        self.kind, self.meta, self.start = kind, meta, None

    def __enter__(self):
        if PROFILER is not None:
            self.start = PROFILER.event()
            self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None:
            end = PROFILER.event()
            end.record()
            PROFILER.record(self.kind, self.meta, self.start, end)
        return False


def set_sm_reserve(n: int) -> None:
    """Keep ``n`` SMs free of the persistent conv grids (room for NCCL under DDP); b2s_set_sm_reserve."""
    _lib.lib().b2s_set_sm_reserve(int(n))


def check(rc: int, what: str = "", launches: int = 1) -> None:
    _check(rc, what)
    STATS["launches"] += launches


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise ValueError(f"libb2s supports float32 and float16 features, got {t.dtype}")


def _cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.B2SError(
                "openpcseg_b200 is a CUDA-only backend: got a CPU tensor "
                "(the reference CPU twin lives in oracle/ and is test infrastructure only)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _i32x3(v: Sequence[int]):
    import ctypes
    return (ctypes.c_int32 * 3)(*[int(x) for x in v])


# --------------------------------------------------------------------------- hashing
def hash_coords(coords: torch.Tensor) -> torch.Tensor:
    _cuda(coords)
    assert coords.dtype == torch.int32 and coords.ndim == 2 and coords.shape[1] == 4, \
        (coords.dtype, coords.shape)
    coords = coords.contiguous()
    out = torch.empty(coords.shape[0], dtype=torch.int64, device=coords.device)
    check(_lib.lib().b2s_hash(coords.data_ptr(), coords.shape[0], out.data_ptr(), _stream()), "hash")
    return out


def kernel_hash(coords: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    _cuda(coords, offsets)
    assert coords.dtype == torch.int32 and coords.ndim == 2 and coords.shape[1] == 4, \
        (coords.dtype, coords.shape)
    assert offsets.dtype == torch.int32 and offsets.ndim == 2 and offsets.shape[1] == 3, \
        (offsets.dtype, offsets.shape)
    coords, offsets = coords.contiguous(), offsets.contiguous()
    n, k = coords.shape[0], offsets.shape[0]
    out = torch.empty((k, n), dtype=torch.int64, device=coords.device)
    check(_lib.lib().b2s_kernel_hash(coords.data_ptr(), n, offsets.data_ptr(), k, out.data_ptr(),
                                     _stream()), "kernel_hash")
    return out


class HashTable:
    """Device hash table {int64 key -> row index}; see b2s_table_* in include/b2s.h."""

    def __init__(self, n: int, device):
        self.n = int(n)
        self.nbytes = _lib.lib().b2s_table_bytes(self.n)
        self.buf = _ws(self.nbytes, device)

    @classmethod
    def from_keys(cls, keys: torch.Tensor) -> "HashTable":
        _cuda(keys)
        keys = keys.contiguous()
        assert keys.dtype == torch.int64 and keys.ndim == 1
        t = cls(keys.shape[0], keys.device)
        check(_lib.lib().b2s_table_build(keys.data_ptr(), t.n, t.buf.data_ptr(), t.nbytes, _stream()),
              "table_build")
        return t

    @classmethod
    def from_coords(cls, coords: torch.Tensor) -> "HashTable":
        _cuda(coords)
        coords = coords.contiguous()
        assert coords.dtype == torch.int32 and coords.ndim == 2 and coords.shape[1] == 4
        t = cls(coords.shape[0], coords.device)
        check(_lib.lib().b2s_table_build_coords(coords.data_ptr(), t.n, t.buf.data_ptr(), t.nbytes,
                                                _stream()), "table_build_coords")
        return t

    def query(self, queries: torch.Tensor) -> torch.Tensor:
        _cuda(queries)
        q = queries.contiguous()
        assert q.dtype == torch.int64
        out = torch.empty(q.shape, dtype=torch.int64, device=q.device)
        check(_lib.lib().b2s_table_query(self.buf.data_ptr(), self.n, q.data_ptr(), q.numel(),
                                         out.data_ptr(), _stream()), "table_query")
        return out


def hash_query(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    """Row index of every query key in ``references`` or -1 (sphashquery semantics)."""
    return HashTable.from_keys(references.view(-1)).query(queries)


def count(idx: torch.Tensor, num: int) -> torch.Tensor:
    _cuda(idx)
    idx = idx.contiguous()
    assert idx.dtype == torch.int32, idx.dtype
    out = torch.empty(int(num), dtype=torch.int32, device=idx.device)
    check(_lib.lib().b2s_count(idx.data_ptr(), idx.numel(), out.data_ptr(), int(num), _stream()),
          "count")
    return out


# ------------------------------------------------------------ unique / downsample
def unique_sorted_i64(keys: torch.Tensor) -> torch.Tensor:
    """Ascending unique of int64 keys (one host sync for the count)."""
    _cuda(keys)
    keys = keys.contiguous().view(-1)
    assert keys.dtype == torch.int64
    n = keys.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=keys.device)
    cnt = torch.empty(2, dtype=torch.int64, device=keys.device)
    nbytes = _lib.lib().b2s_unique_workspace_bytes(n)
    ws = _ws(nbytes, keys.device)
    check(_lib.lib().b2s_unique_i64(keys.data_ptr(), n, out.data_ptr(), cnt.data_ptr(), ws.data_ptr(),
                                    nbytes, _stream()), "unique_i64")
    return out[: int(cnt[0].item())]


def downsample_coords(coords: torch.Tensor, stride, kernel_size, tensor_stride) -> torch.Tensor:
    """Sorted unique output coordinates of a strided conv (spdownsample semantics)."""
    _cuda(coords)
    coords = coords.contiguous()
    assert coords.dtype == torch.int32 and coords.ndim == 2 and coords.shape[1] == 4
    n = coords.shape[0]
    s, k, t = _i32x3(stride), _i32x3(kernel_size), _i32x3(tensor_stride)
    L = _lib.lib()
    cap = L.b2s_downsample_capacity(n, s, k)
    out = torch.empty((max(cap, 1), 4), dtype=torch.int32, device=coords.device)
    cnt = torch.empty(2, dtype=torch.int64, device=coords.device)
    nbytes = L.b2s_downsample_workspace_bytes(n, s, k)
    ws = _ws(nbytes, coords.device)
    check(L.b2s_downsample_coords(coords.data_ptr(), n, s, k, t, out.data_ptr(), cnt.data_ptr(),
                                  ws.data_ptr(), nbytes, _stream()), "downsample_coords")
    m, flag = cnt.tolist()
    if flag:
        raise ValueError("spdownsample: coordinates outside the packable range "
                         "(|x|,|y|,|z| < 2^17 and 0 <= batch < 1024)")
    return out[:m]


# ---------------------------------------------------------------------- kernel map
def kmap_build(in_coords: torch.Tensor, out_coords: torch.Tensor, offsets: torch.Tensor,
               want_nbr_in: bool):
    """(nbr_out int32 [K, N_out], nbr_in int32 [K, N_in] | None, nbsizes int32 [K],
    mask_out int32 [tiles_out, words], mask_in | None) - masks: active offsets per 128-row tile."""
    _cuda(in_coords, out_coords, offsets)
    in_coords, out_coords, offsets = in_coords.contiguous(), out_coords.contiguous(), offsets.contiguous()
    assert in_coords.dtype == out_coords.dtype == offsets.dtype == torch.int32
    n_in, n_out, k = in_coords.shape[0], out_coords.shape[0], offsets.shape[0]
    dev = in_coords.device
    nbr_out = torch.empty((k, n_out), dtype=torch.int32, device=dev)
    nbr_in = torch.empty((k, n_in), dtype=torch.int32, device=dev) if want_nbr_in else None
    nbsizes = torch.empty(k, dtype=torch.int32, device=dev)
    words = (k + 31) // 32
    mask_out = torch.empty((max((n_out + 127) // 128, 1), words), dtype=torch.int32, device=dev)
    mask_in = (torch.empty((max((n_in + 127) // 128, 1), words), dtype=torch.int32, device=dev)
               if want_nbr_in else None)
    L = _lib.lib()
    nbytes = L.b2s_kmap_workspace_bytes(n_in, n_out, k)
    ws = _ws(nbytes, dev)
    check(L.b2s_kmap_build(in_coords.data_ptr(), n_in, out_coords.data_ptr(), n_out,
                           offsets.data_ptr(), k, nbr_out.data_ptr(), _ptr(nbr_in),
                           nbsizes.data_ptr(), mask_out.data_ptr(), _ptr(mask_in), ws.data_ptr(),
                           nbytes, _stream()), "kmap_build")
    return nbr_out, nbr_in, nbsizes, mask_out, mask_in


def tile_mask(nbr: torch.Tensor) -> torch.Tensor:
    """Active-offset bits per 128-row tile of a gather map [K, n] (b2s_tile_mask)."""
    _cuda(nbr)
    nbr = nbr.contiguous()
    k, n = nbr.shape
    mask = torch.empty((max((n + 127) // 128, 1), (k + 31) // 32), dtype=torch.int32, device=nbr.device)
    check(_lib.lib().b2s_tile_mask(nbr.data_ptr(), k, n, mask.data_ptr(), _stream()), "tile_mask")
    return mask


def tile_order_key(nbr: torch.Tensor, nbsizes: torch.Tensor, coords: torch.Tensor, coord_shift: int):
    """(int64 [n] sort keys grouping rows by neighbourhood pattern, uint32-as-int32 [n] presence bits per row)
    (b2s_tile_order_key_bits)."""
    _cuda(nbr, nbsizes, coords)
    k, n = nbr.shape
    keys = torch.empty(n, dtype=torch.int64, device=nbr.device)
    bits = torch.empty(n, dtype=torch.int32, device=nbr.device)
    check(_lib.lib().b2s_tile_order_key_bits(nbr.data_ptr(), k, n, nbsizes.data_ptr(), coords.contiguous().data_ptr(),
                                             int(coord_shift), keys.data_ptr(), bits.data_ptr(), _stream()),
          "tile_order_key")
    return keys, bits


def tile_steps(nbr: torch.Tensor, perm: Optional[torch.Tensor], row_bits: Optional[torch.Tensor], tile_rows: int):
    """Step table of a gather map [K, n] for the tensor-core conv kernel (b2s_tile_steps):
    (tile_mask int32 [tiles, words], step_start int32 [tiles + 1], step_rows int32 [K * tiles * tile_rows],
    tile_rows).  Device-resident; step_rows is allocated at its upper bound (same size as the map)."""
    _cuda(nbr, perm, row_bits)
    assert nbr.dtype == torch.int32 and nbr.is_contiguous()
    k, n = nbr.shape
    tiles = max((n + tile_rows - 1) // tile_rows, 1)
    dev = nbr.device
    mask = torch.empty((tiles, (k + 31) // 32), dtype=torch.int32, device=dev)
    start = torch.empty(tiles + 1, dtype=torch.int32, device=dev)
    rows = torch.empty(k * tiles * tile_rows, dtype=torch.int32, device=dev)
    check(_lib.lib().b2s_tile_steps(nbr.data_ptr(), k, n, _ptr(perm), _ptr(row_bits), int(tile_rows), mask.data_ptr(),
                                    start.data_ptr(), rows.data_ptr(), _stream()), "tile_steps", launches=3)
    return mask, start, rows, int(tile_rows)


def conv_steps_supported(feats: torch.Tensor, c_red: int, c_res: int) -> bool:
    return feats.dtype == torch.float16 and bool(
        _lib.lib().b2s_conv_steps_supported(F16, feats.shape[0], int(c_red), int(c_res)))


def conv_tile_rows(c_res: int, n_rows: int) -> int:
    return int(_lib.lib().b2s_conv_tile_rows(int(c_res), int(n_rows)))


def weight_to_kmajor(w: torch.Tensor) -> torch.Tensor:
    """fp16 [K, C_in, C_out] -> [K, C_out, C_in]: the forward pass's K-major operand (b2s_weight_to_kmajor)."""
    _cuda(w)
    assert w.dtype == torch.float16 and w.ndim == 3 and w.is_contiguous()
    k, c_in, c_out = w.shape
    out = torch.empty((k, c_out, c_in), dtype=torch.float16, device=w.device)
    check(_lib.lib().b2s_weight_to_kmajor(w.data_ptr(), k, c_in, c_out, out.data_ptr(), _stream()), "weight_to_kmajor")
    return out


def weights_refresh(desc: torch.Tensor, total_units: int) -> None:
    """One launch over a device descriptor table int64 [n, 5] = b2s_weight_desc (src, cast, kmajor pointers,
    k | c_in << 32, c_out | unit_start << 32): fp32 masters -> fp16 parameter-layout and K-major copies."""
    _cuda(desc)
    assert desc.dtype == torch.int64 and desc.ndim == 2 and desc.shape[1] == 5 and desc.is_contiguous()
    check(_lib.lib().b2s_weights_refresh(desc.data_ptr(), desc.shape[0], int(total_units), _stream()),
          "weights_refresh")


def kmap_pairs(nbr_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(pairs int32 [K*N_out, 2] padded buffer, d_total int64 [1]) - reference pair order."""
    _cuda(nbr_out)
    k, n_out = nbr_out.shape
    dev = nbr_out.device
    pairs = torch.empty((max(k * n_out, 1), 2), dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    L = _lib.lib()
    nbytes = L.b2s_kmap_workspace_bytes(0, n_out, k)
    ws = _ws(nbytes, dev)
    check(L.b2s_kmap_pairs(nbr_out.data_ptr(), k, n_out, pairs.data_ptr(), total.data_ptr(),
                           ws.data_ptr(), nbytes, _stream()), "kmap_pairs")
    return pairs, total


def kmap_pairs_chunked(nbr_out: torch.Tensor, perm: Optional[torch.Tensor], n_chunks: int):
    """(pairs int32 [K*N_out, 2] padded, seg_sizes int32 [n_chunks * K], d_total int64 [1]): the pair list in
    (row range, offset, row) order for the weight gradient (b2s_kmap_pairs_chunked)."""
    _cuda(nbr_out, perm)
    k, n_out = nbr_out.shape
    dev = nbr_out.device
    pairs = torch.empty((max(k * n_out, 1), 2), dtype=torch.int32, device=dev)
    seg = torch.empty(n_chunks * k, dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    L = _lib.lib()
    nbytes = L.b2s_kmap_pairs_chunked_workspace_bytes(n_out, k, int(n_chunks))
    ws = _ws(nbytes, dev)
    check(L.b2s_kmap_pairs_chunked(nbr_out.data_ptr(), k, n_out, _ptr(perm), int(n_chunks), pairs.data_ptr(),
                                   seg.data_ptr(), total.data_ptr(), ws.data_ptr(), nbytes, _stream()),
          "kmap_pairs_chunked", launches=2)
    return pairs, seg, total


# --------------------------------------------------------------------- convolution
def conv_gather_gemm(feats: torch.Tensor, weight: torch.Tensor, nbr: Optional[torch.Tensor],
                     n_rows: int, transpose_w: bool, flip_k: bool,
                     bias: Optional[torch.Tensor] = None, pairs_hint=None,
                     tile_mask: Optional[torch.Tensor] = None,
                     row_perm: Optional[torch.Tensor] = None, steps=None,
                     weight_kmajor: Optional[torch.Tensor] = None,
                     bn_sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = sum_k feats[nbr[k'][r]] @ (W[k] or W[k]^T); see b2s_conv_gather_gemm(_steps).

    ``steps`` = (tile_mask, step_start, step_rows, tile_rows) of ``tile_steps`` replaces ``nbr`` on the fp16
    tensor-core path; ``weight_kmajor`` = the cached ``weight_to_kmajor(weight)`` (forward only);
    ``bn_sums`` fp64 [2, c_res] (zeroed) receives per-channel sum / sum of squares of the result rows."""
    _cuda(feats, weight, nbr, bias)
    feats, weight = feats.contiguous(), weight.contiguous()
    if weight.ndim == 2:
        weight = weight.unsqueeze(0)
    k, c_in, c_out = weight.shape
    assert weight.dtype == feats.dtype, (weight.dtype, feats.dtype)
    c_red, c_res = (c_out, c_in) if transpose_w else (c_in, c_out)
    if feats.shape[1] != c_red:
        raise ValueError("Input feature size and kernel size mismatch")
    if nbr is not None:
        assert nbr.dtype == torch.int32 and nbr.is_contiguous() and nbr.shape == (k, n_rows), \
            (nbr.dtype, nbr.shape, (k, n_rows))
    if bias is not None:
        bias = bias.to(feats.dtype).contiguous()
    out = torch.empty((n_rows, c_res), dtype=feats.dtype, device=feats.device)
    if n_rows == 0:
        return out
    L = _lib.lib()
    code = _dtype_code(feats)
    w_arg, kmajor = weight, 0
    if transpose_w:
        kmajor = 1                       # the parameter layout [K][c_in][c_out] is this pass's K-major operand
    elif weight_kmajor is not None:
        w_arg, kmajor = weight_kmajor, 1
    nbytes = 0 if kmajor else L.b2s_conv_workspace_bytes(code, n_rows, c_in, c_out, k)
    ws = _ws(nbytes, feats.device) if nbytes else None
    s_mask = s_start = s_rows = None
    tile_rows = 0
    if steps is not None:
        s_mask, s_start, s_rows, tile_rows = steps
        tile_mask = s_mask
    with _Timed("dgrad" if transpose_w else "fwd",
                {"k": k, "c_in": c_in, "c_out": c_out, "rows": n_rows, "dtype": code, "pairs": pairs_hint}):
        check(L.b2s_conv_gather_gemm_steps(code, feats.data_ptr(), feats.shape[0], w_arg.data_ptr(), kmajor, k, c_in,
                                           c_out, int(transpose_w), int(flip_k), _ptr(nbr), _ptr(tile_mask),
                                           _ptr(s_rows), _ptr(s_start), tile_rows, _ptr(row_perm), n_rows,
                                           _ptr(bias), out.data_ptr(), _ptr(bn_sums), _ptr(ws), nbytes, _stream()),
              "conv_gather_gemm", launches=1 if kmajor or code != F16 else 2)
    return out


def conv_wgrad(feats: torch.Tensor, grad_out: torch.Tensor, k: int, pairs: Optional[torch.Tensor],
               nbsizes: Optional[torch.Tensor], swap_pairs: bool, pairs_hint=None) -> torch.Tensor:
    """fp32 grad_w [K, C_in, C_out]; pairs/nbsizes stay on the device (no sync).  ``nbsizes`` with n_chunks * K
    entries = the segment sizes of a ``kmap_pairs_chunked`` list."""
    _cuda(feats, grad_out, pairs, nbsizes)
    feats, grad_out = feats.contiguous(), grad_out.contiguous()
    assert feats.dtype == grad_out.dtype
    c_in, c_out = feats.shape[1], grad_out.shape[1]
    gw = torch.empty((k, c_in, c_out), dtype=torch.float32, device=feats.device)
    with _Timed("wgrad", {"k": k, "c_in": c_in, "c_out": c_out, "rows": feats.shape[0],
                          "dtype": _dtype_code(feats), "pairs": pairs_hint}):
        n_seg = int(nbsizes.numel()) if nbsizes is not None else k
        check(_lib.lib().b2s_conv_wgrad_segments(_dtype_code(feats), feats.data_ptr(), feats.shape[0],
                                                 grad_out.data_ptr(), grad_out.shape[0], k, c_in, c_out,
                                                 _ptr(pairs), _ptr(nbsizes), n_seg, int(swap_pairs),
                                                 gw.data_ptr(), _stream()), "conv_wgrad")
    return gw


# ------------------------------------------------------------------ point <-> voxel
def voxelize_forward(feats: torch.Tensor, idx: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    _cuda(feats, idx, counts)
    feats, idx, counts = feats.contiguous(), idx.contiguous(), counts.contiguous()
    assert idx.dtype == torch.int32 and counts.dtype == torch.int32
    n_pts, c = feats.shape
    n_vox = counts.shape[0]
    code = _dtype_code(feats)
    out = torch.empty((n_vox, c), dtype=feats.dtype, device=feats.device)
    acc = torch.empty((n_vox, c), dtype=torch.float32, device=feats.device) if code == F16 else None
    check(_lib.lib().b2s_voxelize_fwd(code, feats.data_ptr(), idx.data_ptr(), counts.data_ptr(), n_pts,
                                      n_vox, c, out.data_ptr(), _ptr(acc), _stream()), "voxelize_fwd")
    return out


def voxelize_backward(grad_vox: torch.Tensor, idx: torch.Tensor, counts: torch.Tensor,
                      n_pts: int) -> torch.Tensor:
    _cuda(grad_vox, idx, counts)
    grad_vox = grad_vox.contiguous()
    n_vox, c = grad_vox.shape
    out = torch.empty((n_pts, c), dtype=grad_vox.dtype, device=grad_vox.device)
    check(_lib.lib().b2s_voxelize_bwd(_dtype_code(grad_vox), grad_vox.data_ptr(), idx.data_ptr(),
                                      counts.data_ptr(), n_pts, n_vox, c, out.data_ptr(), _stream()),
          "voxelize_bwd")
    return out


def devoxelize_forward(feats: torch.Tensor, idx: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    _cuda(feats, idx, weights)
    feats, idx = feats.contiguous(), idx.contiguous()
    weights = weights.float().contiguous()
    assert idx.dtype == torch.int32 and idx.ndim == 2 and idx.shape[1] == 8, (idx.dtype, idx.shape)
    n_pts, (n_vox, c) = idx.shape[0], feats.shape
    out = torch.empty((n_pts, c), dtype=feats.dtype, device=feats.device)
    check(_lib.lib().b2s_devoxelize_fwd(_dtype_code(feats), feats.data_ptr(), idx.data_ptr(),
                                        weights.data_ptr(), n_pts, n_vox, c, out.data_ptr(), _stream()),
          "devoxelize_fwd")
    return out


def devoxelize_backward(grad_pts: torch.Tensor, idx: torch.Tensor, weights: torch.Tensor,
                        n_vox: int, order: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``order`` int32 [n_pts] (points sorted by idx[:, 0]) selects the run-merging kernel for contended maps."""
    _cuda(grad_pts, idx, weights)
    grad_pts, idx = grad_pts.contiguous(), idx.contiguous()
    weights = weights.float().contiguous()
    n_pts, c = grad_pts.shape
    code = _dtype_code(grad_pts)
    out = torch.empty((n_vox, c), dtype=grad_pts.dtype, device=grad_pts.device)
    acc = torch.empty((n_vox, c), dtype=torch.float32, device=grad_pts.device) if code == F16 else None
    vec = 8 if code == F16 else 4
    if order is not None and c % vec == 0 and c // vec <= 128 and grad_pts.data_ptr() % 16 == 0:
        check(_lib.lib().b2s_devoxelize_bwd_sorted(code, grad_pts.data_ptr(), order.data_ptr(), idx.data_ptr(),
                                                   weights.data_ptr(), n_pts, n_vox, c, out.data_ptr(), _ptr(acc),
                                                   _stream()), "devoxelize_bwd_sorted")
        return out
    check(_lib.lib().b2s_devoxelize_bwd(code, grad_pts.data_ptr(), idx.data_ptr(), weights.data_ptr(),
                                        n_pts, n_vox, c, out.data_ptr(), _ptr(acc), _stream()),
          "devoxelize_bwd")
    return out


def scatter_max(feats: torch.Tensor, idx: torch.Tensor, m: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(out [m, c], arg int64 [m, c]) like torch_scatter.scatter_max(feats, idx, dim=0, dim_size=m)."""
    _cuda(feats, idx)
    feats, idx = feats.contiguous(), idx.long().contiguous()
    n, c = feats.shape
    out = torch.empty((m, c), dtype=feats.dtype, device=feats.device)
    arg = torch.empty((m, c), dtype=torch.int64, device=feats.device)
    keys = torch.empty((m, c), dtype=torch.int32, device=feats.device)
    check(_lib.lib().b2s_scatter_max(_dtype_code(feats), feats.data_ptr(), idx.data_ptr(), n, c, int(m),
                                     out.data_ptr(), arg.data_ptr(), keys.data_ptr(), _stream()),
          "scatter_max", launches=3)
    return out, arg


def trilinear_map(pts: torch.Tensor, vox_coords: torch.Tensor, stride: int,
                  table: Optional[HashTable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(idx int32 [N, 8], weights fp32 [N, 8]) of voxel_to_point at ``stride`` (fused)."""
    _cuda(pts, vox_coords)
    pts = pts.contiguous()
    assert pts.dtype == torch.float32 and pts.ndim == 2 and pts.shape[1] == 4, (pts.dtype, pts.shape)
    if table is None:
        table = HashTable.from_coords(vox_coords)
    n = pts.shape[0]
    idx = torch.empty((n, 8), dtype=torch.int32, device=pts.device)
    w = torch.empty((n, 8), dtype=torch.float32, device=pts.device)
    check(_lib.lib().b2s_trilinear_map(pts.data_ptr(), n, int(stride), table.buf.data_ptr(), table.n,
                                       idx.data_ptr(), w.data_ptr(), _stream()), "trilinear_map")
    return idx, w


def ti_weights(pts: torch.Tensor, idx_query: torch.Tensor, scale: float) -> torch.Tensor:
    """calc_ti_weights: idx_query int64 [8, N] -> fp32 [8, N]."""
    _cuda(pts, idx_query)
    pts = pts.float().contiguous()
    if pts.shape[1] == 3:
        pts = torch.cat([pts, torch.zeros_like(pts[:, :1])], 1)
    idx_query = idx_query.long().contiguous()
    n = pts.shape[0]
    assert idx_query.shape == (8, n), idx_query.shape
    w = torch.empty((8, n), dtype=torch.float32, device=pts.device)
    check(_lib.lib().b2s_ti_weights(pts.data_ptr(), n, idx_query.data_ptr(), float(scale), w.data_ptr(),
                                    _stream()), "ti_weights")
    return w


# ---------------------------------------------------------------- fused batch norm
def bn_supported(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype in (torch.float16, torch.float32) and x.ndim == 2 and x.shape[0] > 0
            and bool(_lib.lib().b2s_bn_supported(_dtype_code(x), x.shape[1])))


def bn_forward(x: torch.Tensor, residual: Optional[torch.Tensor], gamma, beta, running_mean, running_var,
               eps: float, momentum: float, relu: bool, sums: Optional[torch.Tensor] = None):
    """y = act(bn_train(x) [+ residual]); returns (y, mean, invstd).  ``sums`` fp64 [2, c]: per-channel sum and
    sum of squares of x already accumulated by the producing conv (conv_gather_gemm(bn_sums=...))."""
    _cuda(x, residual, gamma, beta, running_mean, running_var)
    x = x.contiguous()
    if residual is not None:
        residual = residual.contiguous()
        assert residual.shape == x.shape and residual.dtype == x.dtype
    n, c = x.shape
    y = torch.empty_like(x)
    stat = torch.empty((4, c), dtype=torch.float32, device=x.device)     # mean, invstd, scale, shift
    ready = sums is not None
    if not ready:
        sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    check(_lib.lib().b2s_bn_forward_sums(_dtype_code(x), x.data_ptr(), _ptr(residual), n, c, _ptr(gamma),
                                         _ptr(beta), float(eps), float(momentum), _ptr(running_mean),
                                         _ptr(running_var), int(relu), y.data_ptr(), stat[0].data_ptr(),
                                         stat[1].data_ptr(), stat[2].data_ptr(), sums.data_ptr(), int(ready),
                                         _stream()),
          "bn_forward", launches=2 if ready else 3)
    return y, stat[0], stat[1], stat[2:4]


def bn_stats(x: torch.Tensor, extra: int = 0) -> torch.Tensor:
    """fp64 [2*c + extra]: per-channel sum and sum of squares of the rows of x (b2s_bn_stats)."""
    _cuda(x)
    x = x.contiguous()
    n, c = x.shape
    sums = torch.zeros(2 * c + extra, dtype=torch.float64, device=x.device)
    check(_lib.lib().b2s_bn_stats(_dtype_code(x), x.data_ptr(), n, c, sums.data_ptr(), _stream()), "bn_stats", launches=1)
    return sums


def bn_forward_global(x, residual, gamma, beta, running_mean, running_var, eps, momentum, relu, sums_count):
    """bn_forward from ALL-REDUCED statistics: ``sums_count`` fp64 [2c + 1] = global sums and the global row count."""
    _cuda(x, residual)
    x = x.contiguous()
    if residual is not None:
        residual = residual.contiguous()
    n, c = x.shape
    y = torch.empty_like(x)
    stat = torch.empty((4, c), dtype=torch.float32, device=x.device)
    check(_lib.lib().b2s_bn_forward_sums(_dtype_code(x), x.data_ptr(), _ptr(residual), n, c, _ptr(gamma),
                                         _ptr(beta), float(eps), float(momentum), _ptr(running_mean),
                                         _ptr(running_var), int(relu), y.data_ptr(), stat[0].data_ptr(),
                                         stat[1].data_ptr(), stat[2].data_ptr(), sums_count.data_ptr(), 2,
                                         _stream()), "bn_forward", launches=2)
    return y, stat[0], stat[1], stat[2:4]


def _relu_mode(relu: bool, y, scale_shift) -> int:
    """0 none / 1 mask from the saved output / 2 mask recomputed from x (no residual: y was not kept)."""
    if not relu:
        return 0
    return 1 if y is not None else 2


def bn_backward_reduce(dy, y, x, mean, invstd, relu: bool, scale_shift=None) -> torch.Tensor:
    """fp64 [2, c] = (sum dy', sum dy' * xhat) over the local rows (dy' = dy masked by the ReLU)."""
    _cuda(dy, y, x)
    dy = dy.contiguous()
    n, c = x.shape
    sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    check(_lib.lib().b2s_bn_backward_reduce(_dtype_code(x), dy.data_ptr(), _ptr(y), x.data_ptr(), n, c,
                                            mean.data_ptr(), invstd.data_ptr(), _relu_mode(relu, y, scale_shift),
                                            _ptr(scale_shift), sums.data_ptr(), _stream()), "bn_backward_reduce")
    return sums


def bn_backward_apply(dy, y, x, mean, invstd, gamma, relu: bool, want_dres: bool, sums, n_total, scale_shift=None):
    """(dx, dres | None) from (all-reduced) ``sums`` and the global row count ``n_total`` (device fp64 [1])."""
    _cuda(dy, y, x)
    dy = dy.contiguous()
    n, c = x.shape
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    check(_lib.lib().b2s_bn_backward_apply(_dtype_code(x), dy.data_ptr(), _ptr(y), x.data_ptr(), n, c,
                                           mean.data_ptr(), invstd.data_ptr(), _ptr(gamma),
                                           _relu_mode(relu, y, scale_shift), _ptr(scale_shift), dx.data_ptr(),
                                           _ptr(dres), sums.data_ptr(), _ptr(n_total), _stream()),
          "bn_backward_apply")
    return dx, dres


def bn_backward(dy: torch.Tensor, y: Optional[torch.Tensor], x: torch.Tensor, mean, invstd, gamma,
                relu: bool, want_dres: bool, scale_shift=None):
    """Returns (dx, dres | None, d_gamma fp32 [c], d_beta fp32 [c]).  ``y`` None with ``relu``: the mask is
    recomputed from x and ``scale_shift`` (no-residual layers do not keep their output for backward)."""
    sums = bn_backward_reduce(dy, y, x, mean, invstd, relu, scale_shift)
    dx, dres = bn_backward_apply(dy, y, x, mean, invstd, gamma, relu, want_dres, sums, None, scale_shift)
    STATS["launches"] += 0
    return dx, dres, sums[1].float(), sums[0].float()


# ------------------------------------------------------------------- range-image ops
def map_count(pxpy: torch.Tensor, b: int, h: int, w: int) -> torch.Tensor:
    _cuda(pxpy)
    pxpy = pxpy.contiguous()
    assert pxpy.dtype == torch.int32 and pxpy.shape[1] == 3
    out = torch.empty((b, 1, h, w), dtype=torch.int32, device=pxpy.device)
    check(_lib.lib().b2s_map_count(pxpy.data_ptr(), pxpy.shape[0], b, h, w, out.data_ptr(), _stream()),
          "map_count")
    return out


def denselize_forward(feats: torch.Tensor, count_map: torch.Tensor, pxpy: torch.Tensor) -> torch.Tensor:
    _cuda(feats, count_map, pxpy)
    feats, count_map, pxpy = feats.float().contiguous(), count_map.contiguous(), pxpy.contiguous()
    b, _, h, w = count_map.shape
    n, c = feats.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=feats.device)
    check(_lib.lib().b2s_denselize_fwd(feats.data_ptr(), pxpy.data_ptr(), count_map.data_ptr(), n, c, b,
                                       h, w, out.data_ptr(), _stream()), "denselize_fwd")
    return out


def denselize_backward(grad: torch.Tensor, count_map: torch.Tensor, pxpy: torch.Tensor) -> torch.Tensor:
    _cuda(grad, count_map, pxpy)
    grad, count_map, pxpy = grad.float().contiguous(), count_map.contiguous(), pxpy.contiguous()
    b, c, h, w = grad.shape
    n = pxpy.shape[0]
    out = torch.empty((n, c), dtype=torch.float32, device=grad.device)
    check(_lib.lib().b2s_denselize_bwd(grad.data_ptr(), pxpy.data_ptr(), count_map.data_ptr(), n, c, b,
                                       h, w, out.data_ptr(), _stream()), "denselize_bwd")
    return out


# ------------------------------------------- reference pybind names (drop-in boundary)
def hash_cuda(idx):                                   # TS/backend/hash/hash_cuda.cu:67-73
    return hash_coords(idx)


def kernel_hash_cuda(idx, kernel_offset):             # hash_cuda.cu:75-84
    return kernel_hash(idx, kernel_offset)


def hash_query_cuda(hash_query_, hash_target, idx_target):   # TS/backend/others/query_cuda.cu:9-56
    """Returns idx_target[row] + 1 for hits and 0 for misses, like the reference."""
    res = hash_query(hash_query_, hash_target)
    hit = res >= 0
    out = torch.zeros_like(res)
    out[hit] = idx_target[res[hit]] + 1
    return out


def count_cuda(idx, s):                               # TS/backend/others/count_cuda.cu:25-31
    return count(idx, s)


def voxelize_forward_cuda(inputs, idx, counts):       # TS/backend/voxelize/voxelize_cuda.cu:44-61
    return voxelize_forward(inputs, idx, counts)


def voxelize_backward_cuda(top_grad, idx, counts, n):  # voxelize_cuda.cu:63-80
    return voxelize_backward(top_grad, idx, counts, n)


def devoxelize_forward_cuda(feat, indices, weight):   # TS/backend/devoxelize/devoxelize_cuda.cu:61-78
    return devoxelize_forward(feat, indices, weight)


def devoxelize_backward_cuda(top_grad, indices, weight, n):  # devoxelize_cuda.cu:82-98
    return devoxelize_backward(top_grad, indices, weight, n)


def _nbr_from_pairs(nbmap: torch.Tensor, nbsizes_host: torch.Tensor, n_rows: int, col_src: int,
                    col_dst: int) -> torch.Tensor:
    """Pair list (reference format) -> gather map [K, n_rows]; boundary adapter only."""
    k = nbsizes_host.numel()
    nbr = torch.full((k, n_rows), -1, dtype=torch.int32, device=nbmap.device)
    kk = torch.repeat_interleave(torch.arange(k, device=nbmap.device),
                                 nbsizes_host.to(nbmap.device).long())
    nbr[kk, nbmap[:, col_dst].long()] = nbmap[:, col_src].int()
    return nbr


def convolution_forward_cuda(in_feat, out_feat, kernel, neighbor_map, neighbor_offset, transpose):
    """TS/backend/convolution/convolution_cuda.cu:53-165 signature: writes ``out_feat`` in place."""
    n_rows = out_feat.shape[0]
    nbr = _nbr_from_pairs(neighbor_map, neighbor_offset, n_rows, 1 if transpose else 0,
                          0 if transpose else 1)
    out_feat.copy_(conv_gather_gemm(in_feat, kernel, nbr, n_rows, False, False))


def convolution_backward_cuda(in_feat, grad_in_feat, grad_out_feat, kernel, grad_kernel, neighbor_map,
                              neighbor_offset, transpose):
    """convolution_cuda.cu:167-278 signature: fills ``grad_in_feat`` and ``grad_kernel``."""
    n_in = in_feat.shape[0]
    nbr = _nbr_from_pairs(neighbor_map, neighbor_offset, n_in, 0 if transpose else 1,
                          1 if transpose else 0)
    grad_in_feat.resize_as_(in_feat).copy_(
        conv_gather_gemm(grad_out_feat.contiguous(), kernel, nbr, n_in, True, False))
    gw = conv_wgrad(in_feat, grad_out_feat.contiguous(), kernel.shape[0],
                    neighbor_map.int().contiguous(),
                    neighbor_offset.int().to(in_feat.device).contiguous(), bool(transpose))
    grad_kernel.resize_as_(kernel).copy_(gw.to(kernel.dtype))
