"""Deterministic synthetic SemanticKITTI-shaped scans.

HDL-64E-like: 64 elevation angles linspace(+2 deg, -24.8 deg) x 1875 azimuths = 120 000
rays; range = min(ground plane 1.73 m below the sensor, one wall per 10-degree sector at
U(8, 45) m, 80 m) + N(0, 0.02 m); intensity U(0, 1).  Pre-processing mirrors the voxel
dataset of the reference (pcseg/data/dataset/semantickitti/semantickitti_voxel.py:112-129):
``pc = round(xyz / voxel)``, shift to non-negative, keep one point per voxel.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .torchsparse.utils.quantize import sparse_quantize

__all__ = ["make_scan", "make_raw_scan", "make_batch", "make_model_batch"]


def _rays(seed: int, n_beams: int, n_azimuth: int):
    """(rng, xyz fp32 [R, 3], intensity fp32 [R, 1], ring int [R]) of one synthetic sweep, beam-major."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))[:, None]
    azim = np.linspace(-np.pi, np.pi, n_azimuth, endpoint=False)[None, :]
    sector = ((azim + np.pi) / np.deg2rad(10.0)).astype(np.int64) % 36
    wall = rng.uniform(8.0, 45.0, size=36)[sector]                      # [1, A] horizontal distance
    with np.errstate(divide="ignore"):
        r_ground = np.where(np.sin(elev) < 0, 1.73 / -np.sin(elev), np.inf)
    r = np.minimum(np.minimum(r_ground, wall / np.cos(elev)), 80.0)
    r = r + rng.normal(0.0, 0.02, size=r.shape)
    xyz = np.stack([r * np.cos(elev) * np.cos(azim), r * np.cos(elev) * np.sin(azim),
                    r * np.sin(elev)], -1).reshape(-1, 3).astype(np.float32)
    intensity = rng.uniform(0.0, 1.0, size=(xyz.shape[0], 1)).astype(np.float32)
    ring = np.repeat(np.arange(n_beams), n_azimuth)
    return rng, xyz, intensity, ring


def make_scan(seed: int = 0, voxel_size: float = 0.05, n_beams: int = 64, n_azimuth: int = 1875,
              num_class: int = 20) -> Dict[str, np.ndarray]:
    rng, xyz, intensity, ring = _rays(seed, n_beams, n_azimuth)
    pc = np.round(xyz / voxel_size).astype(np.int32)
    pc -= pc.min(0, keepdims=True)
    _, keep = sparse_quantize(pc, 1, return_index=True)
    feats = np.concatenate([xyz, intensity], 1)[keep]
    labels = rng.integers(0, num_class, size=keep.shape[0]).astype(np.int64)
    return {"coords": pc[keep].astype(np.int32), "feats": feats.astype(np.float32), "labels": labels,
            "ring": ring[keep].astype(np.float32), "n_raw_points": int(xyz.shape[0])}


def make_raw_scan(seed: int = 0, n_beams: int = 64, n_azimuth: int = 1875, num_class: int = 20):
    """Every ray of the sweep (no voxel de-duplication): points fp32 [R, 5] = (x, y, z, intensity, ring)
    and per-point labels - the input of the cylinder front-end (semantickitti_cylinder.py:144-171)."""
    _, xyz, intensity, ring = _rays(seed, n_beams, n_azimuth)
    labels = np.random.default_rng(seed + 7919).integers(0, num_class, size=xyz.shape[0]).astype(np.int64)
    return {"points": np.concatenate([xyz, intensity, ring[:, None].astype(np.float32)], 1), "labels": labels}


def make_batch(seeds: List[int], **kw) -> Dict[str, np.ndarray]:
    """Collate scans: coords int32 [N, 4] = (x, y, z, batch), feats fp32 [N, 4], labels int64 [N]."""
    scans = [make_scan(s, **kw) for s in seeds]
    coords = np.concatenate([np.concatenate([s["coords"], np.full((len(s["coords"]), 1), b, np.int32)], 1)
                             for b, s in enumerate(scans)])
    return {"coords": coords, "feats": np.concatenate([s["feats"] for s in scans]),
            "labels": np.concatenate([s["labels"] for s in scans]),
            "n_scans": len(scans), "n_raw_points": sum(s["n_raw_points"] for s in scans)}


def make_model_batch(kind: str, seeds: List[int], **kw) -> Dict[str, np.ndarray]:
    """Host arrays of one batch for the reference's four sparse segmentors,
    collated like the reference's dataset classes:

    ``voxel``     MinkUNet / SPVCNN: coords int32 [N,4], feats fp32 [N,4], labels, offset
                  (semantickitti_voxel.py:112-141 + collate)
    ``fusion``    RPVNet: feats fp32 [N,5] = (x,y,z,intensity,ring), range_image fp32 [B,5,64,2048],
                  range_pxpy fp32 [N,3] = (batch, px, py) in [-1,1] (semantickitti_fusion.py:64-114,199-220),
                  no random yaw cut
    ``cylinder``  Cylinder3D: point_feature fp32 [P,9], point_coord int64 [P,4] = (rho,phi,z cell, batch),
                  voxel_coord int64 [V,4], voxel_label, point_label, offset
                  (semantickitti_cylinder.py:144-171,176-200; grid 480x360x32 over rho [0,50], phi [-180,180],
                  z [-4,2] as cylinder_cy480_cr10.yaml)
    """
    import torch
    from . import frontend as FE
    if kind in ("voxel", "fusion"):
        b = make_batch(seeds, **kw)
        n_per = np.bincount(b["coords"][:, 3], minlength=len(seeds))
        out = {"coords": b["coords"], "feats": b["feats"], "labels": b["labels"],
               "offset": np.cumsum(n_per).astype(np.int32), "n_scans": len(seeds)}
        if kind == "fusion":
            scans = [make_scan(s, **kw) for s in seeds]
            feats5 = [np.concatenate([s["feats"], s["ring"][:, None]], 1).astype(np.float32) for s in scans]
            imgs, pxpy = [], []
            for i, f in enumerate(feats5):
                img, pp = FE.range_projection(torch.from_numpy(f), 0.0, (kw.get("n_beams", 64), 2048))
                imgs.append(img.numpy())
                pxpy.append(np.concatenate([np.full((len(f), 1), i, np.float64), pp.numpy()], 1))
            out["feats"] = np.concatenate(feats5)
            out["range_image"] = np.stack(imgs).astype(np.float32)
            out["range_pxpy"] = np.concatenate(pxpy).astype(np.float32)
        return out
    assert kind == "cylinder", kind
    keys = ("point_feature", "point_coord", "point_label", "voxel_coord", "voxel_label")
    parts = {k: [] for k in keys}
    offset = []
    for i, s in enumerate(seeds):
        raw = make_raw_scan(s, **{k: v for k, v in kw.items() if k in ("n_beams", "n_azimuth", "num_class")})
        d = FE.cylinder_scan(torch.from_numpy(raw["points"][:, :4]), torch.from_numpy(raw["labels"]),
                             (480, 360, 32), (0.0, -180.0, -4.0), (50.0, 180.0, 2.0), kw.get("num_class", 20))
        for k in keys:
            v = d[k].numpy()
            if k.endswith("_coord"):
                v = np.concatenate([v.astype(np.int64), np.full((len(v), 1), i, np.int64)], 1)
            parts[k].append(v)
        offset.append(len(d["voxel_coord"]))
    out = {k: np.concatenate(v) for k, v in parts.items()}
    out["offset"] = np.cumsum(offset).astype(np.int32)
    out["n_scans"] = len(seeds)
    return out
