"""Host-side voxel quantisation of a raw scan (reference: TS/utils/quantize.py:10-46)."""
from typing import Tuple, Union

import numpy as np

__all__ = ["sparse_quantize", "ravel_hash"]


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
    vox = np.floor(coords / np.asarray(voxel_size)).astype(np.int32)
    _, first, inverse = np.unique(ravel_hash(vox), return_index=True, return_inverse=True)
    out = [vox[first]]
    if return_index:
        out.append(first)
    if return_inverse:
        out.append(inverse)
    return out[0] if len(out) == 1 else out
