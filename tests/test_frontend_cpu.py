"""Tensor-program data front-end (openpcseg_b200/frontend.py) against the numpy restatement of the
reference datasets (oracle/ref_frontend.py).  CPU tensors here; the same code runs on CUDA tensors with
the device sparse_quantize that tests/test_gpu_frontend.py pins."""
import numpy as np
import torch

from openpcseg_b200 import frontend
from openpcseg_b200.synthetic import make_scan
from oracle import ref_frontend as R


def _scan(seed, n=20000):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.normal(0, 12, (n, 2)), rng.normal(-1, 1.2, (n, 1)), rng.uniform(0, 1, (n, 1))], 1)
    labels = rng.integers(0, 20, n)
    labels[rng.random(n) < 0.05] = 67
    return pts, labels


def test_voxel_scan_matches_reference_dataset_code():
    for dtype in (np.float32, np.float64):
        pts, labels = _scan(0)
        pts = pts.astype(dtype)
        ref = R.voxel_scan_ref(pts.copy(), labels, 0.05)
        got = frontend.voxel_scan(torch.from_numpy(pts), torch.from_numpy(labels), 0.05)
        np.testing.assert_array_equal(got["lidar"].C.numpy(), ref["pc"])
        np.testing.assert_array_equal(got["lidar"].F.numpy(), ref["feat"])
        np.testing.assert_array_equal(got["targets"].F.numpy(), ref["labels"])
        np.testing.assert_array_equal(got["inverse_map"].F.numpy(), ref["inverse_map"])
        np.testing.assert_array_equal(got["targets_mapped"].C.numpy(), ref["pc_all"])


def test_voxelize_with_label_majority_and_first_maximum():
    rng = np.random.default_rng(1)
    coords = rng.integers(0, 12, (5000, 3))
    labels = rng.integers(0, 5, 5000)
    labels[::7] = 67
    vc, vl, inds, inv = R.voxelize_with_label_ref(coords, labels, 20)
    gc, gl, gi, gn = frontend.voxelize_with_label(torch.from_numpy(coords), torch.from_numpy(labels), 20)
    np.testing.assert_array_equal(gc.numpy(), vc)
    np.testing.assert_array_equal(gl.numpy(), vl)          # ties resolve to the smallest class, like argmax
    np.testing.assert_array_equal(gi.numpy(), inds)
    np.testing.assert_array_equal(gn.numpy(), inv)
    # a cell whose points are all label 67 gets class 0
    c2 = np.array([[0, 0, 0], [0, 0, 0], [1, 1, 1]])
    l2 = np.array([67, 67, 3])
    assert frontend.voxelize_with_label(torch.from_numpy(c2), torch.from_numpy(l2), 20)[1].tolist() == [0, 3]


def test_cylinder_scan_matches_reference_dataset_code():
    pts, labels = _scan(2)
    args = dict(grid_size=[480, 360, 32], min_bound=[0.0, -180.0, -4.0], max_bound=[50.0, 180.0, 2.0], num_classes=20)
    ref = R.cylinder_scan_ref(pts.copy(), labels, **args)
    got = frontend.cylinder_scan(torch.from_numpy(pts), torch.from_numpy(labels), **args)
    for key in ("voxel_coord", "voxel_label", "inverse_map", "point_coord"):
        np.testing.assert_array_equal(got[key].numpy(), ref[key], err_msg=key)
    for key in ("voxel_feature", "point_feature"):
        np.testing.assert_allclose(got[key].numpy(), ref[key], rtol=0, atol=1e-5, err_msg=key)
    assert got["voxel_feature"].shape[1] == 9 and got["point_feature"].shape[1] == 9


def test_range_projection_last_point_wins():
    rng = np.random.default_rng(3)
    n = 30000
    pts = np.concatenate([rng.normal(0, 15, (n, 2)), rng.normal(-1, 1, (n, 1)), rng.uniform(0, 1, (n, 1)),
                          rng.integers(0, 64, (n, 1)).astype(np.float64)], 1)
    img_ref, pxpy_ref = R.range_projection_ref(pts, 0.37)
    img, pxpy = frontend.range_projection(torch.from_numpy(pts), 0.37)
    np.testing.assert_array_equal(pxpy.numpy(), pxpy_ref)
    np.testing.assert_array_equal(img.numpy(), img_ref)    # collisions exist: ~n^2 / (2 * 64 * 2048) of them


def test_front_end_reproduces_the_benchmark_scan():
    scan = make_scan(0)                       # already one point per voxel, in sparse_quantize order
    out = frontend.voxel_scan(torch.from_numpy(scan["feats"]), torch.from_numpy(scan["labels"]), 0.05)
    np.testing.assert_array_equal(out["lidar"].C.numpy(), scan["coords"])
    np.testing.assert_array_equal(out["lidar"].F.numpy(), scan["feats"])
    np.testing.assert_array_equal(out["inverse_map"].F.numpy(), np.arange(scan["coords"].shape[0]))


# ----------------------------------------------------- pinned to outputs of the reference's own functions
def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend.npz"))


def test_sparse_quantize_matches_reference_output():
    from openpcseg_b200.torchsparse.utils.quantize import sparse_quantize
    g = _golden()
    vox, idx, inv = sparse_quantize(g["q_pts"], 0.25, return_index=True, return_inverse=True)
    np.testing.assert_array_equal(vox, g["q_vox"])
    np.testing.assert_array_equal(idx, g["q_idx"])
    np.testing.assert_array_equal(inv, g["q_inv"])


def test_cylinder_pieces_match_reference_output():
    g = _golden()
    vc, vl, inds, inv = R.voxelize_with_label_ref(g["c_coords"], g["c_labels"], 20)          # the restatement
    for got, key in ((vc, "c_vcoords"), (vl, "c_vlabels"), (inds, "c_inds"), (inv, "c_inverse")):
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    gc, gl, gi, gn = frontend.voxelize_with_label(torch.from_numpy(g["c_coords"]), torch.from_numpy(g["c_labels"]), 20)
    for got, key in ((gc, "c_vcoords"), (gl, "c_vlabels"), (gi, "c_inds"), (gn, "c_inverse")):
        np.testing.assert_array_equal(got.numpy(), g[key], err_msg=key)
    # cart2polar (radians) against the polar part of the tensor program (degrees)
    args = dict(grid_size=[480, 360, 32], min_bound=[0.0, -180.0, -4.0], max_bound=[50.0, 180.0, 2.0], num_classes=20)
    out = frontend.cylinder_scan(torch.from_numpy(g["c_scan"]), torch.from_numpy(g["c_labels"]), **args)
    pol = g["c_polar"].copy()
    pol[:, 1] = pol[:, 1] / np.pi * 180.0
    np.testing.assert_allclose(out["point_feature"][:, 3:6].numpy(), pol.astype(np.float32), rtol=0, atol=2e-5)


def test_range_projection_matches_reference_output():
    g = _golden()
    img, pxpy = frontend.range_projection(torch.from_numpy(g["r_points"]), float(g["r_yaw_offset"]))
    np.testing.assert_array_equal(pxpy.numpy(), g["r_pxpy"])
    np.testing.assert_array_equal(img.numpy(), g["r_image"])
    img_r, pxpy_r = R.range_projection_ref(g["r_points"], float(g["r_yaw_offset"]))
    np.testing.assert_array_equal(img_r, g["r_image"])
    np.testing.assert_array_equal(pxpy_r, g["r_pxpy"])
