"""Point <-> voxel glue of the voxel / fusion segmentors, on fused device ops.

Same contracts as pcseg/model/segmentor/voxel/minkunet/utils.py:11-105 (and its copies
under fusion/spvcnn, fusion/rpvnet): cache keys, voxel order (= ascending hash after
``initial_voxelize``), corner order and weights of ``voxel_to_point``.  Implementation
differences: one fused kernel builds the 8-corner index + trilinear weights
(b2s_trilinear_map), the sorted-unique of the hashes is a device radix sort, and voxel
coordinates are always averaged in fp32 (the reference averages them in fp16 under AMP,
which is inexact beyond 2048 cells - SURVEY.md section 5).
"""
from __future__ import annotations

import torch

from .. import backend as B
from ..torchsparse import PointTensor, SparseTensor
from ..torchsparse.nn import functional as F

__all__ = ["initial_voxelize", "point_to_voxel", "voxel_to_point"]


def _grid_coords(z: PointTensor, stride: int) -> torch.Tensor:
    """floor(point / stride) * stride with the batch column, int32 [N, 4]."""
    xyz = torch.floor(z.C[:, :3] / stride).int() * stride
    return torch.cat([xyz, z.C[:, -1].int().view(-1, 1)], 1)


def initial_voxelize(z: PointTensor, init_res: float, after_res: float) -> SparseTensor:
    new_float_coord = torch.cat([(z.C[:, :3] * init_res) / after_res, z.C[:, -1].view(-1, 1)], 1)
    floored = torch.floor(new_float_coord)
    pc_hash = F.sphash(floored.int())
    sparse_hash = B.unique_sorted_i64(pc_hash)            # ascending => voxel order
    idx_query = F.sphashquery(pc_hash, sparse_hash)
    idx32 = idx_query.int()
    counts = F.spcount(idx32, sparse_hash.shape[0])
    with torch.autocast("cuda", enabled=False):           # coordinates are never fp16
        coords = torch.round(F.spvoxelize(floored.float(), idx32, counts)).int()
    feats = F.spvoxelize(z.F, idx32, counts)
    x = SparseTensor(feats, coords, 1)
    x.cmaps.setdefault(x.stride, x.coords)
    z.additional_features["idx_query"][1] = idx_query
    z.additional_features["counts"][1] = counts
    z.C = new_float_coord
    return x


def point_to_voxel(x: SparseTensor, z: PointTensor) -> SparseTensor:
    cache_i, cache_c = z.additional_features["idx_query"], z.additional_features["counts"]
    if cache_i.get(x.s) is None:
        idx_query = B.HashTable.from_coords(x.C).query(F.sphash(_grid_coords(z, x.s[0])))
        cache_i[x.s] = idx_query
        cache_c[x.s] = F.spcount(idx_query.int(), x.C.shape[0])
    out = SparseTensor(F.spvoxelize(z.F, cache_i[x.s], cache_c[x.s]), x.C, x.s)
    out.cmaps, out.kmaps = x.cmaps, x.kmaps
    return out


def voxel_to_point(x: SparseTensor, z: PointTensor, nearest: bool = False) -> PointTensor:
    if z.idx_query.get(x.s) is None or z.weights.get(x.s) is None:
        idx_query, weights = B.trilinear_map(z.C.float(), x.C, x.s[0])
        if nearest:
            weights[:, 1:] = 0.0
            idx_query[:, 1:] = -1
        z.idx_query[x.s], z.weights[x.s] = idx_query, weights
    out = PointTensor(F.spdevoxelize(x.F, z.idx_query[x.s], z.weights[x.s]), z.C,
                      idx_query=z.idx_query, weights=z.weights)
    out.additional_features = z.additional_features
    return out
