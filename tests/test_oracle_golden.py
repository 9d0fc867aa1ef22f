"""Pin the CPU oracle (oracle/ref_ops.py) against outputs of the reference itself.

The fixtures in tests/golden were produced by tests/golden/make_golden.py running
the reference's python package + compiled CPU backend.  Integer results must match
bit-for-bit; floating-point results within 1e-5 relative (fp32).
"""
import numpy as np
import pytest

from oracle import ref_ops as R


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() /
                 max(float(np.abs(b).max()), 1e-30))


def test_hash_known_answers(golden):
    g = golden("hash_offsets")
    assert np.array_equal(R.sphash(g["ka_coords"]), g["ka_hash"])
    # the vectors quoted in SURVEY.md section 8(c)
    assert R.sphash(np.array([[0, 0, 0, 0]], np.int32))[0] == 947293587111810033
    assert R.sphash(np.array([[7, 7, 7, 1]], np.int32))[0] == 137767048215775683
    assert R.sphash(np.array([[2147483647, -2147483648, 0, 3]], np.int32))[0] == 533978114288266694
    assert np.array_equal(R.sphash(g["rand_coords"]), g["rand_hash"])


@pytest.mark.parametrize("name,ks,st", [("k3", 3, 1), ("k2s4", 2, 4), ("k133", (1, 3, 3), 1),
                                         ("k313", (3, 1, 3), 2), ("k311", (3, 1, 1), 1),
                                         ("k3s8", 3, 8)])
def test_offsets_and_kernel_hash(golden, name, ks, st):
    g = golden("hash_offsets")
    off = R.get_kernel_offsets(ks, st)
    assert np.array_equal(off, g[f"off_{name}"])
    assert np.array_equal(R.sphash(g["rand_coords"], off), g[f"khash_{name}"])


def test_five_voxel_known_answer(golden):
    g = golden("five_voxel")
    c = g["coords"]
    nb, ns = R.build_kmap(c, c, 3)
    assert np.array_equal(nb, g["k3_nbmaps"]) and np.array_equal(ns, g["k3_nbsizes"])
    assert ns.tolist() == [1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 5, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 1]
    f = np.arange(5, dtype=np.float32)[:, None]
    out = R.conv_forward(f, np.ones((27, 1, 1), np.float32), nb, ns, (5, 5))
    assert np.array_equal(out, g["k3_out"]) and out.ravel().tolist() == [6, 6, 6, 6, 4]
    oc = R.spdownsample(c, 2, 2, 1)
    assert np.array_equal(oc, g["k2s2_coords"])
    nb2, ns2 = R.build_kmap(c, oc, 2)
    assert np.array_equal(nb2, g["k2s2_nbmaps"]) and np.array_equal(ns2, g["k2s2_nbsizes"])
    o2 = R.conv_forward(f, np.ones((8, 1, 1), np.float32), nb2, ns2, (5, oc.shape[0]))
    assert np.array_equal(o2, g["k2s2_out"])
    o3 = R.conv_forward(o2, np.ones((8, 1, 1), np.float32), nb2, ns2, (5, oc.shape[0]), transposed=True)
    assert np.array_equal(o3, g["k2s2t_out"])
    assert np.array_equal(g["k2s2t_coords"], c)


@pytest.mark.parametrize("tag,ks", [("k3", 3), ("k133", (1, 3, 3)), ("k313", (3, 1, 3)),
                                     ("k311", (3, 1, 1))])
def test_submanifold_maps_and_conv(golden, tag, ks):
    g = golden("conv_maps")
    c, x = g["coords"], g["feats"]
    nb, ns = R.build_kmap(c, c, ks)
    assert np.array_equal(nb, g[f"{tag}_nbmaps"]) and np.array_equal(ns, g[f"{tag}_nbsizes"])
    w = g[f"{tag}_w"]
    out = R.conv_forward(x, w, nb, ns, (c.shape[0], c.shape[0]))
    assert rel_err(out, g[f"{tag}_out"]) < 1e-5
    gin, gw = R.conv_backward(x, w, g[f"{tag}_gout"], nb, ns)
    assert rel_err(gin, g[f"{tag}_gin"]) < 1e-5 and rel_err(gw, g[f"{tag}_gw"]) < 1e-5


@pytest.mark.parametrize("tag,stride", [("k2s2", (2, 2, 2)), ("k3s2", (2, 2, 2)),
                                         ("k3s221", (2, 2, 1))])
def test_strided_maps_and_conv(golden, tag, stride):
    g = golden("conv_maps")
    c, x = g["coords"], g["feats"]
    ks = 2 if tag == "k2s2" else 3
    oc = R.spdownsample(c, stride, ks, 1)
    assert np.array_equal(oc, g[f"{tag}_coords"])
    nb, ns = R.build_kmap(c, oc, ks)
    assert np.array_equal(nb, g[f"{tag}_nbmaps"]) and np.array_equal(ns, g[f"{tag}_nbsizes"])
    w = g[f"{tag}_w"]
    out = R.conv_forward(x, w, nb, ns, (c.shape[0], oc.shape[0]))
    assert rel_err(out, g[f"{tag}_out"]) < 1e-5
    gin, gw = R.conv_backward(x, w, g[f"{tag}_gout"], nb, ns)
    assert rel_err(gin, g[f"{tag}_gin"]) < 1e-5 and rel_err(gw, g[f"{tag}_gw"]) < 1e-5


def test_coarse_level_and_transposed(golden):
    g = golden("conv_maps")
    c1 = g["k2s2_coords"]
    nb, ns = R.build_kmap(c1, c1, 3, in_stride=2)
    assert np.array_equal(nb, g["s2k3_nbmaps"]) and np.array_equal(ns, g["s2k3_nbsizes"])
    x1 = g["k2s2_out"]
    out = R.conv_forward(x1, g["s2k3_w"], nb, ns, (c1.shape[0], c1.shape[0]))
    assert rel_err(out, g["s2k3_out"]) < 1e-5
    # transposed k2s2 reuses the forward map with swapped columns
    nb2, ns2 = g["k2s2_nbmaps"], g["k2s2_nbsizes"]
    n0 = g["coords"].shape[0]
    out_t = R.conv_forward(x1, g["k2s2t_w"], nb2, ns2, (n0, c1.shape[0]), transposed=True)
    assert rel_err(out_t, g["k2s2t_out"]) < 1e-5
    assert np.array_equal(g["k2s2t_coords"], g["coords"])
    gin, gw = R.conv_backward(x1, g["k2s2t_w"], g["k2s2t_gout"], nb2, ns2, transposed=True)
    assert rel_err(gin, g["k2s2t_gin"]) < 1e-5 and rel_err(gw, g["k2s2t_gw"]) < 1e-5


@pytest.mark.parametrize("tag,st,ks,tst", [("s2k2", 2, 2, 1), ("s2k2_t2", 2, 2, 2), ("s2k3", 2, 3, 1),
                                            ("s221k3", (2, 2, 1), 3, 1), ("s2k3_t2", 2, 3, 2)])
def test_downsample(golden, tag, st, ks, tst):
    g = golden("downsample")
    assert np.array_equal(R.spdownsample(g[f"{tag}_in"], st, ks, tst), g[f"{tag}_out"])


def test_point_voxel_pipeline(golden):
    g = golden("point_voxel")
    vc, vf, idx, cnt, nfc = R.initial_voxelize(g["pts"], g["pt_feats"], 0.05, 0.05)
    assert np.array_equal(vc, g["iv_coords"])
    assert np.array_equal(idx, g["iv_idx_query"]) and np.array_equal(cnt, g["iv_counts"])
    assert rel_err(vf, g["iv_feats"]) < 1e-5
    i1, w1 = R.trilinear_map(nfc, vc, 1)
    assert np.array_equal(i1, g["v2p1_idx"]) and rel_err(w1, g["v2p1_w"]) < 1e-5
    assert rel_err(R.spdevoxelize_forward(g["v2p1_vfeats"], i1, w1), g["v2p1_out"]) < 1e-5
    i2, w2 = R.trilinear_map(nfc, g["s2_coords"], 2)
    assert np.array_equal(i2, g["v2p2_idx"]) and rel_err(w2, g["v2p2_w"]) < 1e-5
    assert rel_err(R.spdevoxelize_forward(g["s2_feats"], i2, w2), g["v2p2_out"]) < 1e-5
    pi, pc = R.point_to_voxel_map(nfc, g["s2_coords"], 2)
    assert np.array_equal(pi, g["p2v2_idx"]) and np.array_equal(pc, g["p2v2_counts"])
    assert rel_err(R.spvoxelize_forward(g["pt_feats"], pi, pc), g["p2v2_out"]) < 1e-5
    gb = R.spvoxelize_backward(g["vox_bwd_gout"], pi, pc, g["pts"].shape[0])
    assert rel_err(gb, g["vox_bwd_gin"]) < 1e-5


def test_devoxelize_backward_is_adjoint():
    """The reference CPU twin of devoxelize-backward is wrong (devoxelize_cpu.cpp:48-53),
    so the oracle is checked as the exact adjoint of the (golden-pinned) forward."""
    rng = np.random.default_rng(3)
    n_vox, n_pts, c = 50, 200, 5
    idx = rng.integers(-1, n_vox, size=(n_pts, 8))
    w = rng.random((n_pts, 8)).astype(np.float32)
    f = rng.standard_normal((n_vox, c)).astype(np.float32)
    gp = rng.standard_normal((n_pts, c)).astype(np.float32)
    lhs = float((R.spdevoxelize_forward(f, idx, w).astype(np.float64) * gp).sum())
    rhs = float((R.spdevoxelize_backward(gp, idx, w, n_vox).astype(np.float64) * f).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
