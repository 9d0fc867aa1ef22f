"""Voxel quantisation of a raw scan (reference: TS/utils/quantize.py:10-46).

numpy input -> the reference's host path (np.unique of the ravel hash); CUDA tensor input -> the same
result (same voxel order, same first-occurrence indices, same inverse map) computed on the device with
the backend's sort/unique and hash table, so the data front-end needs no CPU pass.
"""
from typing import Tuple, Union

import numpy as np
import torch

__all__ = ["sparse_quantize", "sparse_quantize_device", "ravel_hash"]


def ravel_hash(x: np.ndarray) -> np.ndarray:
    """Mixed-radix key of integer rows: unique per distinct row inside the bounding box."""
    assert x.ndim == 2, x.shape
    x = (x - x.min(axis=0)).astype(np.uint64)
    radix = x.max(axis=0).astype(np.uint64) + np.uint64(1)
    key = np.zeros(x.shape[0], dtype=np.uint64)
    for d in range(x.shape[1]):
        key = key * radix[d] + x[:, d]
    return key


def sparse_quantize(coords, voxel_size: Union[float, Tuple[float, ...]] = 1, *,
                    return_index: bool = False, return_inverse: bool = False):
    if isinstance(voxel_size, (float, int)):
        voxel_size = (voxel_size,) * 3
    assert isinstance(voxel_size, tuple) and len(voxel_size) == 3
    if isinstance(coords, torch.Tensor) and coords.is_cuda:
        return sparse_quantize_device(coords, voxel_size, return_index=return_index,
                                      return_inverse=return_inverse)
    vox = np.floor(coords / np.asarray(voxel_size)).astype(np.int32)
    _, first, inverse = np.unique(ravel_hash(vox), return_index=True, return_inverse=True)
    out = [vox[first]]
    if return_index:
        out.append(first)
    if return_inverse:
        out.append(inverse)
    return out[0] if len(out) == 1 else out


def sparse_quantize_device(coords: torch.Tensor, voxel_size=(1, 1, 1), *, return_index: bool = False,
                           return_inverse: bool = False):
    """Device twin of ``sparse_quantize``: coords float/int ``[N, 3]`` CUDA tensor.

    Voxels come out in ascending ravel-key order (x, then y, then z after the min-shift), ``index``
    is the FIRST point of every voxel in input order and ``inverse`` maps points to voxels - exactly
    what np.unique(return_index, return_inverse) yields on the host path."""
    from ... import backend as B

    assert coords.is_cuda and coords.dim() == 2 and coords.shape[1] == 3, coords.shape
    size = torch.as_tensor(voxel_size, dtype=torch.float64, device=coords.device)
    vox = torch.floor(coords.double() / size).to(torch.int32)
    if vox.shape[0] == 0:
        empty = torch.empty(0, dtype=torch.int64, device=coords.device)
        out = [vox] + ([empty] if return_index else []) + ([empty] if return_inverse else [])
        return out[0] if len(out) == 1 else out
    rel = (vox - vox.min(dim=0).values).long()
    radix = rel.max(dim=0).values + 1
    key = (rel[:, 0] * radix[1] + rel[:, 1]) * radix[2] + rel[:, 2]     # ravel_hash, same digit order
    uniq = B.unique_sorted_i64(key)
    # duplicate keys resolve to the smallest row (atomicMin in b2s_table_build) = first occurrence
    first = B.HashTable.from_keys(key).query(uniq)
    out = [vox.index_select(0, first)]
    if return_index:
        out.append(first)
    if return_inverse:
        out.append(B.HashTable.from_keys(uniq).query(key))
    return out[0] if len(out) == 1 else out
