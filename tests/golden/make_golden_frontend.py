#!/usr/bin/env python
"""Freeze outputs of the reference's own per-scan preprocessing (build container only):
torchsparse.utils.quantize.sparse_quantize (zip), and cart2polar / voxelize_with_label /
SemkittiFusionDataset.get_range_image of pcseg/data/dataset/semantickitti.  Modules the datasets import
but this path never calls (SharedArray, torch_scatter, cv2, ...) are stubbed.  Nothing of the reference
is copied; inputs and outputs are stored.   python tests/golden/make_golden_frontend.py -> frontend.npz"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG                                             # noqa: E402


def import_with_stubs(name, tries=20):
    for _ in range(tries):
        try:
            return importlib.import_module(name)
        except ModuleNotFoundError as exc:
            missing = exc.name
            if missing is None or missing.startswith("pcseg") or missing.startswith("tools"):
                raise
            stub = types.ModuleType(missing)
            stub.__path__ = []                                       # behaves as a package for sub-imports
            sys.modules[missing] = stub
            print("stubbed", missing)
    raise RuntimeError("too many missing modules")


def main():
    tmp = tempfile.mkdtemp()
    MG.import_reference(tmp)
    from torchsparse.utils.quantize import sparse_quantize           # the reference's, from the zip
    sys.path.insert(0, "/root/reference")
    if not hasattr(np, "int"):
        np.int = int                                                 # the reference targets numpy < 1.24
    cyl = import_with_stubs("pcseg.data.dataset.semantickitti.semantickitti_cylinder")
    fus = import_with_stubs("pcseg.data.dataset.semantickitti.semantickitti_fusion")
    rng = np.random.default_rng(0)
    out = {}
    # ---- sparse_quantize
    pts = rng.uniform(-20, 20, size=(5000, 3))
    pts[:500] = pts[2500:3000]
    vox, idx, inv = sparse_quantize(pts, 0.25, return_index=True, return_inverse=True)
    out.update(q_pts=pts, q_vox=vox, q_idx=idx, q_inv=inv)
    # ---- cylinder: polar transform + majority labels
    n = 8000
    scan = np.concatenate([rng.normal(0, 12, (n, 2)), rng.normal(-1, 1.2, (n, 1)), rng.uniform(0, 1, (n, 1))], 1)
    labels = rng.integers(0, 20, n)
    labels[rng.random(n) < 0.05] = 67
    pol = cyl.cart2polar(scan[:, :3])
    coords = rng.integers(0, 14, size=(n, 3))
    vc, vl, vi, vn = cyl.voxelize_with_label(coords, labels, 20)
    out.update(c_scan=scan, c_labels=labels, c_polar=pol, c_coords=coords, c_vcoords=vc, c_vlabels=vl, c_inds=vi,
               c_inverse=vn)
    # ---- range projection (the random cut is drawn from np.random: seed it and record the draw)
    m = 20000
    p5 = np.concatenate([rng.normal(0, 15, (m, 2)), rng.normal(-1, 1, (m, 1)), rng.uniform(0, 1, (m, 1)),
                         rng.integers(0, 64, (m, 1)).astype(np.float64)], 1)
    import cv2 as cv2_stub
    if not hasattr(cv2_stub, "resize"):                              # INIT_HW == UP_HW: the resize is the identity
        cv2_stub.resize = lambda img, size, interpolation=None: img.copy()
        cv2_stub.INTER_LINEAR = 1
    np.random.seed(123)
    draw = np.random.rand()
    np.random.seed(123)
    ds = fus.SemkittiFusionDataset.__new__(fus.SemkittiFusionDataset)
    image, pxpy = fus.SemkittiFusionDataset.get_range_image(ds, p5)
    out.update(r_points=p5, r_yaw_offset=np.array((draw - 0.5) * 2 * np.pi), r_image=image, r_pxpy=pxpy)
    np.savez_compressed(os.path.join(HERE, "frontend.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
