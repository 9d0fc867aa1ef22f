"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures produced
by the reference and against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): integer / index results bit-exact; fp32 activations and
gradients <= 1e-5 relative (max-norm), fp16 <= 1e-3.
"""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
FP16_TOL = 1e-3


@pytest.fixture(scope="module")
def ts():
    import openpcseg_b200.torchsparse as ts_mod
    assert torch.cuda.is_available()
    return ts_mod


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def rel_err(a, b):
    a = a.detach().float().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else a.astype(np.float64)
    b = b.detach().float().cpu().numpy().astype(np.float64) if torch.is_tensor(b) else b.astype(np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def eq(t, a):
    return np.array_equal(t.cpu().numpy(), a)


def multi_batch_cloud(seed, n=4000, extent=60, batches=3):
    rng = np.random.default_rng(seed)
    out = []
    for b in range(batches):
        xy = rng.integers(0, extent, size=(n, 2))
        z = np.where(rng.random(n) < 0.7, (xy[:, 0] // 5 + xy[:, 1] // 7) % 4, rng.integers(0, 20, n))
        c = np.unique(np.stack([xy[:, 0], xy[:, 1], z], 1), axis=0)
        rng.shuffle(c)
        out.append(np.concatenate([c, np.full((len(c), 1), b)], 1))
    return np.concatenate(out).astype(np.int32)


# ------------------------------------------------------------------------------ hashing
def test_hash_golden_and_oracle(ts, golden):
    F = ts.nn.functional
    g = golden("hash_offsets")
    assert eq(F.sphash(dev(g["ka_coords"])), g["ka_hash"])
    assert eq(F.sphash(dev(g["rand_coords"])), g["rand_hash"])
    for name in ["k3", "k2s4", "k133", "k313", "k311", "k3s8"]:
        off = dev(g[f"off_{name}"])
        assert eq(F.sphash(dev(g["rand_coords"]), off), g[f"khash_{name}"])
    c = multi_batch_cloud(1)
    c[::7, :3] -= 40                                     # negatives
    assert eq(F.sphash(dev(c)), R.sphash(c))
    off = R.get_kernel_offsets(3, 2)
    assert eq(F.sphash(dev(c), dev(off)), R.sphash(c, off))
    assert F.sphash(torch.zeros((0, 4), dtype=torch.int32, device="cuda")).shape == (0,)


def test_kernel_offsets_match_oracle(ts):
    from openpcseg_b200.torchsparse.nn.utils import get_kernel_offsets
    for ks, st, dil in [(3, 1, 1), (2, 2, 1), ((1, 3, 3), 4, 1), ((3, 1, 3), (2, 2, 1), 1), (5, 1, 1),
                        ((3, 1, 1), 1, 2), (1, 1, 1), (4, 1, 1)]:
        assert np.array_equal(get_kernel_offsets(ks, st, dil).numpy(), R.get_kernel_offsets(ks, st, dil))


def test_hash_query_and_count(ts):
    F = ts.nn.functional
    rng = np.random.default_rng(5)
    refs = rng.integers(0, 1 << 59, size=5000).astype(np.int64)
    refs[100] = refs[7]                                  # duplicate key: first row wins
    q = np.concatenate([refs[rng.integers(0, 5000, 3000)], rng.integers(0, 1 << 59, 3000)]).astype(np.int64)
    q = q.reshape(6, 1000)
    got = F.sphashquery(dev(q), dev(refs))
    assert got.shape == (6, 1000) and eq(got, R.sphashquery(q, refs))
    idx = rng.integers(-1, 50, size=10000).astype(np.int32)
    assert eq(F.spcount(dev(idx), 50), R.spcount(idx, 50))
    assert F.sphashquery(torch.zeros(0, dtype=torch.int64, device="cuda"), dev(refs)).numel() == 0
    none = F.sphashquery(dev(q), torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert bool((none == -1).all())


def test_reference_backend_names(ts):
    """The pybind-level names of the reference (TS/backend/pybind_cuda.cpp:18-39)."""
    from openpcseg_b200 import backend as B
    rng = np.random.default_rng(2)
    refs = np.unique(rng.integers(0, 1 << 50, 300)).astype(np.int64)
    q = np.concatenate([refs[:50], np.array([3, 5], np.int64)])
    idx_target = torch.arange(len(refs), device="cuda")
    out = B.hash_query_cuda(dev(q), dev(refs), idx_target).cpu().numpy()
    exp = R.sphashquery(q, refs) + 1
    assert np.array_equal(out, exp)
    c = multi_batch_cloud(3, n=500, extent=20, batches=1)
    assert eq(B.hash_cuda(dev(c)), R.sphash(c))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.float16, FP16_TOL)])
@pytest.mark.parametrize("case", ["k3s1", "k2s2", "k2s2_transposed"])
def test_reference_backend_conv_names(ts, case, dtype, tol):
    """``backend.convolution_{forward,backward}_cuda`` with the REFERENCE's argument format (pair list
    ``nbmaps`` [M, 2] = (in, out) grouped by offset + host ``nbsizes``; TS/backend/convolution/
    convolution_cuda.cu:53-278, called from TS/nn/functional/conv.py:80-137) against the oracle."""
    from openpcseg_b200 import backend as B
    rng = np.random.default_rng(31)
    c = multi_batch_cloud(9, n=1500, extent=40, batches=2)
    c_in, c_out = 32, 64
    if case == "k3s1":
        lo, hi, ks, transposed = c, c, 3, False
    else:
        lo, hi, ks = c, R.spdownsample(c, 2, 2, 1), 2
        transposed = case.endswith("transposed")
    nbmaps, nbsizes = R.build_kmap(lo, hi, ks, 1)                 # (lo row, hi row) pairs
    n_in, n_out = (len(hi), len(lo)) if transposed else (len(lo), len(hi))
    x = rng.standard_normal((n_in, c_in)).astype(np.float32)
    w = (rng.standard_normal((ks ** 3, c_in, c_out)) / np.sqrt(c_in * ks)).astype(np.float32)
    g = rng.standard_normal((n_out, c_out)).astype(np.float32)
    if dtype == torch.float16:                                    # the oracle sees the rounded operands
        x, w, g = (a.astype(np.float16).astype(np.float32) for a in (x, w, g))
    exp_y = R.conv_forward(x, w, nbmaps, nbsizes, (len(lo), len(hi)), transposed)
    exp_gx, exp_gw = R.conv_backward(x, w, g, nbmaps, nbsizes, transposed)
    nb_d, sz_h = dev(nbmaps.astype(np.int32)), torch.from_numpy(nbsizes.astype(np.int32))
    xd, wd, gd = dev(x, dtype), dev(w, dtype), dev(g, dtype)
    y = torch.zeros(n_out, c_out, device="cuda", dtype=dtype)
    B.convolution_forward_cuda(xd, y, wd, nb_d, sz_h, transposed)
    assert rel_err(y, exp_y) <= tol
    gx, gw = torch.zeros_like(xd), torch.zeros_like(wd)
    B.convolution_backward_cuda(xd, gx, gd, wd, gw, nb_d, sz_h, transposed)
    assert rel_err(gx, exp_gx) <= tol
    assert rel_err(gw, exp_gw) <= (tol if dtype == torch.float32 else 2 * tol)


# --------------------------------------------------------------- coordinates and maps
def test_unique_and_downsample(ts, golden):
    from openpcseg_b200 import backend as B
    F = ts.nn.functional
    rng = np.random.default_rng(11)
    keys = rng.integers(-(1 << 40), 1 << 59, size=20000).astype(np.int64)
    keys[5000:10000] = keys[:5000]
    assert eq(B.unique_sorted_i64(dev(keys)), np.unique(keys))
    g = golden("downsample")
    for tag, (st, ks, tst) in {"s2k2": (2, 2, 1), "s2k2_t2": (2, 2, 2), "s2k3": (2, 3, 1),
                               "s221k3": ((2, 2, 1), 3, 1), "s2k3_t2": (2, 3, 2)}.items():
        assert eq(F.spdownsample(dev(g[f"{tag}_in"]), st, ks, tst), g[f"{tag}_out"]), tag
    c = multi_batch_cloud(4)
    for st, ks, tst in [(2, 2, 1), (2, 3, 1), ((2, 2, 1), 3, 1), (2, 2, 4)]:
        src = c if tst == 1 else R.spdownsample(R.spdownsample(c, 2, 2, 1), 2, 2, 2)
        assert eq(F.spdownsample(dev(src), st, ks, tst), R.spdownsample(src, st, ks, tst)), (st, ks, tst)


@pytest.mark.parametrize("tag,ks", [("k3", 3), ("k133", (1, 3, 3)), ("k313", (3, 1, 3)), ("k311", (3, 1, 1))])
def test_kmap_golden_submanifold(ts, golden, tag, ks):
    F = ts.nn.functional
    g = golden("conv_maps")
    c = dev(g["coords"])
    km = F.build_kernel_map(c, c, ks, (1, 1, 1))
    nbmaps, nbsizes, sizes = km
    assert eq(nbmaps, g[f"{tag}_nbmaps"]) and eq(nbsizes, g[f"{tag}_nbsizes"])
    assert sizes == (c.shape[0], c.shape[0])
    # symmetric maps: nbr_in[k] == nbr_out[K-1-k]
    _, nbr_in, _, mask_out, mask_in = __import__("openpcseg_b200").backend.kmap_build(
        c, c, ts.nn.utils.get_kernel_offsets(ks, 1, 1, "cuda"), True)
    assert torch.equal(nbr_in, km.nbr_out.flip(0))
    # tile masks: bit k of tile t <=> some row of the tile has a neighbour for offset k
    for nbr, mask in ((km.nbr_out, mask_out), (nbr_in, mask_in)):
        kv, n = nbr.shape
        pad = (-n) % 128
        have = torch.nn.functional.pad(nbr >= 0, (0, pad)).view(kv, -1, 128).any(2)      # [K, tiles]
        bits = (have.t().long() << torch.arange(kv, device="cuda")).sum(1)
        assert torch.equal(bits, mask[:, 0].long() & 0xFFFFFFFF)


def test_kmap_multibatch_and_strided(ts, golden):
    F = ts.nn.functional
    g = golden("conv_maps")
    c = dev(g["coords"])
    for tag, st, ks in [("k2s2", 2, 2), ("k3s2", 2, 3), ("k3s221", (2, 2, 1), 3)]:
        oc = F.spdownsample(c, st, ks, 1)
        assert eq(oc, g[f"{tag}_coords"])
        nbmaps, nbsizes, _ = F.build_kernel_map(c, oc, ks, (1, 1, 1))
        assert eq(nbmaps, g[f"{tag}_nbmaps"]) and eq(nbsizes, g[f"{tag}_nbsizes"]), tag
    cm = multi_batch_cloud(6)
    for ks, in_stride in [(3, 1), (2, 1), ((1, 3, 3), 1)]:
        oc = cm if ks != 2 else R.spdownsample(cm, 2, 2, 1)
        nb, ns = R.build_kmap(cm, oc, ks, in_stride)
        km = F.build_kernel_map(dev(cm), dev(cm) if ks != 2 else dev(oc), ks, (in_stride,) * 3)
        assert eq(km[0], nb) and eq(km[1], ns)
        # per-offset injectivity + nbr_in is the inverse of nbr_out
        no = km.nbr_out.cpu().numpy()
        for k in range(no.shape[0]):
            v = no[k][no[k] >= 0]
            assert len(np.unique(v)) == len(v)


# ------------------------------------------------------------------------ convolution
def _sparse(ts, feats, coords, stride=1):
    x = ts.SparseTensor(feats, coords, stride)
    x.cmaps[x.stride] = x.coords
    return x


@pytest.mark.parametrize("tag,ks,stride", [("k3", 3, 1), ("k133", (1, 3, 3), 1), ("k311", (3, 1, 1), 1),
                                            ("k2s2", 2, 2), ("k3s2", 3, 2), ("k3s221", 3, (2, 2, 1))])
def test_conv_fp32_golden(ts, golden, tag, ks, stride):
    F = ts.nn.functional
    g = golden("conv_maps")
    x = dev(g["feats"]).requires_grad_(True)
    w = dev(g[f"{tag}_w"]).requires_grad_(True)
    y = F.conv3d(_sparse(ts, x, dev(g["coords"])), w, ks, stride=stride)
    assert eq(y.coords, g[f"{tag}_coords"])
    assert rel_err(y.feats, g[f"{tag}_out"]) < FP32_TOL
    y.feats.backward(dev(g[f"{tag}_gout"]))
    assert rel_err(x.grad, g[f"{tag}_gin"]) < FP32_TOL
    assert rel_err(w.grad, g[f"{tag}_gw"]) < FP32_TOL


def test_conv_transposed_and_level2_golden(ts, golden):
    F = ts.nn.functional
    g = golden("conv_maps")
    x0 = _sparse(ts, dev(g["feats"]), dev(g["coords"]))
    x1 = F.conv3d(x0, dev(g["k2s2_w"]), 2, stride=2)
    assert rel_err(x1.feats, g["k2s2_out"]) < FP32_TOL
    # k3 at stride 2 reuses cmaps[(2,2,2)]
    x1i = ts.SparseTensor(dev(g["k2s2_out"]).requires_grad_(True), x1.coords, x1.stride)
    x1i.cmaps, x1i.kmaps = x1.cmaps, x1.kmaps
    w = dev(g["s2k3_w"]).requires_grad_(True)
    y = F.conv3d(x1i, w, 3)
    assert rel_err(y.feats, g["s2k3_out"]) < FP32_TOL
    nbmaps, nbsizes, _ = x1.kmaps[((2, 2, 2), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    assert eq(nbmaps, g["s2k3_nbmaps"]) and eq(nbsizes, g["s2k3_nbsizes"])
    y.feats.backward(dev(g["s2k3_gout"]))
    assert rel_err(x1i.feats.grad, g["s2k3_gin"]) < FP32_TOL and rel_err(w.grad, g["s2k3_gw"]) < FP32_TOL
    # transposed k2s2 back to the stride-1 coordinates
    xt = ts.SparseTensor(dev(g["k2s2_out"]).requires_grad_(True), x1.coords, x1.stride)
    xt.cmaps, xt.kmaps = x1.cmaps, x1.kmaps
    wt = dev(g["k2s2t_w"]).requires_grad_(True)
    yt = F.conv3d(xt, wt, 2, stride=2, transposed=True)
    assert eq(yt.coords, g["coords"]) and yt.stride == (1, 1, 1)
    assert rel_err(yt.feats, g["k2s2t_out"]) < FP32_TOL
    yt.feats.backward(dev(g["k2s2t_gout"]))
    assert rel_err(xt.feats.grad, g["k2s2t_gin"]) < FP32_TOL and rel_err(wt.grad, g["k2s2t_gw"]) < FP32_TOL


@pytest.mark.parametrize("cin,cout", [(4, 32), (32, 32), (64, 96), (96, 96), (128, 128), (192, 128), (20, 20)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_conv_vs_oracle_channels(ts, cin, cout, dtype):
    """Submanifold k3 + strided k2s2 + transposed, multi-batch, channel shapes of MinkUNet."""
    F = ts.nn.functional
    tol = FP32_TOL if dtype == torch.float32 else FP16_TOL
    c = multi_batch_cloud(21, n=1500, extent=30, batches=2)
    rng = np.random.default_rng(cin * 131 + cout)
    x = rng.standard_normal((c.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    if dtype == torch.float16:                           # compare on the fp16-rounded operands
        x = x.astype(np.float16).astype(np.float32)
        w = w.astype(np.float16).astype(np.float32)
    nb, ns = R.build_kmap(c, c, 3)
    go = rng.standard_normal((c.shape[0], cout)).astype(np.float32)
    if dtype == torch.float16:
        go = go.astype(np.float16).astype(np.float32)
    exp = R.conv_forward(x, w, nb, ns, (len(c), len(c)))
    egi, egw = R.conv_backward(x, w, go, nb, ns)
    xt = dev(x, dtype).requires_grad_(True)
    wt = dev(w).requires_grad_(True)                      # fp32 master weights
    with torch.autocast("cuda", dtype=torch.float16, enabled=dtype == torch.float16):
        y = F.conv3d(_sparse(ts, xt, dev(c)), wt, 3)
    assert y.feats.dtype == dtype
    assert rel_err(y.feats, exp) < tol
    y.feats.backward(dev(go, dtype))
    assert rel_err(xt.grad, egi) < tol
    assert wt.grad.dtype == torch.float32 and rel_err(wt.grad, egw) < tol


def test_conv_1x1_bias_and_errors(ts):
    F = ts.nn.functional
    c = multi_batch_cloud(8, n=600, extent=20, batches=2)
    x = torch.randn(len(c), 24, device="cuda", requires_grad=True)
    w = torch.randn(24, 40, device="cuda", requires_grad=True)
    b = torch.randn(40, device="cuda")
    y = F.conv3d(_sparse(ts, x, dev(c)), w, 1, bias=b)
    ref = x.detach() @ w.detach() + b
    assert rel_err(y.feats, ref) < FP32_TOL
    go = torch.randn_like(y.feats)
    y.feats.backward(go)
    assert rel_err(x.grad, go @ w.detach().t()) < FP32_TOL
    assert rel_err(w.grad, x.detach().t() @ go) < FP32_TOL
    with pytest.raises(ValueError):                       # channel mismatch, like the reference
        F.conv3d(_sparse(ts, torch.randn(len(c), 5, device="cuda"), dev(c)), torch.randn(27, 6, 8, device="cuda"), 3)
    # empty tensor
    e = _sparse(ts, torch.zeros(0, 8, device="cuda"), torch.zeros(0, 4, dtype=torch.int32, device="cuda"))
    assert F.conv3d(e, torch.randn(27, 8, 8, device="cuda"), 3).feats.shape == (0, 8)


def test_module_surface(ts):
    spnn = ts.nn
    c = multi_batch_cloud(9, n=800, extent=24, batches=2)
    net = torch.nn.Sequential(spnn.Conv3d(4, 16, 3), spnn.BatchNorm(16), spnn.ReLU(True),
                              spnn.Conv3d(16, 16, 2, stride=2), spnn.BatchNorm(16), spnn.ReLU(True),
                              spnn.Conv3d(16, 8, 2, stride=2, transposed=True)).cuda()
    x = _sparse(ts, torch.randn(len(c), 4, device="cuda"), dev(c))
    y = net(x)
    assert y.feats.shape == (len(c), 8) and y.stride == (1, 1, 1)
    y.feats.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    z = ts.cat([y, x])
    assert z.feats.shape[1] == 12 and z.kmaps is x.kmaps


# -------------------------------------------------------------------- point <-> voxel
def test_point_voxel_golden(ts, golden):
    from openpcseg_b200 import backend as B
    F = ts.nn.functional
    g = golden("point_voxel")
    idx = dev(g["p2v2_idx"]).int()
    cnt = dev(g["p2v2_counts"])
    assert rel_err(F.spvoxelize(dev(g["pt_feats"]), idx, cnt), g["p2v2_out"]) < FP32_TOL
    assert rel_err(B.voxelize_backward(dev(g["vox_bwd_gout"]), idx, cnt, g["pts"].shape[0]),
                   g["vox_bwd_gin"]) < FP32_TOL
    for lvl, vc, vf in [("1", g["iv_coords"], g["v2p1_vfeats"]), ("2", g["s2_coords"], g["s2_feats"])]:
        nfc = R.initial_voxelize(g["pts"], g["pt_feats"], 0.05, 0.05)[4]
        ii, ww = B.trilinear_map(dev(nfc), dev(vc), int(lvl))
        assert eq(ii, g[f"v2p{lvl}_idx"]) and rel_err(ww, g[f"v2p{lvl}_w"]) < FP32_TOL
        out = F.spdevoxelize(dev(vf), ii, ww)
        assert rel_err(out, g[f"v2p{lvl}_out"]) < FP32_TOL
        # calc_ti_weights in the reference layout
        w8 = F.calc_ti_weights(dev(nfc), dev(g[f"v2p{lvl}_idx"]).t().contiguous(), scale=int(lvl))
        assert rel_err(w8.t(), g[f"v2p{lvl}_w"]) < FP32_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("c", [4, 32, 96, 10])
def test_voxelize_devoxelize_vs_oracle(ts, dtype, c):
    F = ts.nn.functional
    tol = FP32_TOL if dtype == torch.float32 else FP16_TOL
    rng = np.random.default_rng(c)
    n_pts, n_vox = 5000, 700
    idx = rng.integers(-1, n_vox, size=n_pts).astype(np.int32)
    cnt = R.spcount(idx, n_vox)
    f = rng.standard_normal((n_pts, c)).astype(np.float32)
    if dtype == torch.float16:
        f = f.astype(np.float16).astype(np.float32)
    ft = dev(f, dtype).requires_grad_(True)
    out = F.spvoxelize(ft, dev(idx), dev(cnt))
    assert out.dtype == dtype and rel_err(out, R.spvoxelize_forward(f, idx, cnt)) < tol
    g = rng.standard_normal((n_vox, c)).astype(np.float32)
    if dtype == torch.float16:
        g = g.astype(np.float16).astype(np.float32)
    out.backward(dev(g, dtype))
    assert rel_err(ft.grad, R.spvoxelize_backward(g, idx, cnt, n_pts)) < tol
    # devoxelize
    i8 = rng.integers(-1, n_vox, size=(n_pts, 8)).astype(np.int32)
    w8 = rng.random((n_pts, 8)).astype(np.float32)
    vf = rng.standard_normal((n_vox, c)).astype(np.float32)
    gp = rng.standard_normal((n_pts, c)).astype(np.float32)
    if dtype == torch.float16:
        vf = vf.astype(np.float16).astype(np.float32)
        gp = gp.astype(np.float16).astype(np.float32)
    vt = dev(vf, dtype).requires_grad_(True)
    o = F.spdevoxelize(vt, dev(i8), dev(w8))
    assert rel_err(o, R.spdevoxelize_forward(vf, i8, w8)) < tol
    o.backward(dev(gp, dtype))
    assert rel_err(vt.grad, R.spdevoxelize_backward(gp, i8, w8, n_vox)) < tol


def test_range_image_ops():
    from openpcseg_b200 import backend as B
    rng = np.random.default_rng(0)
    n, c, b, h, w = 3000, 6, 2, 8, 32
    pxpy = np.stack([rng.integers(0, b, n), rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.int32)
    feats = rng.standard_normal((n, c)).astype(np.float32)
    cm = B.map_count(dev(pxpy), b, h, w)
    exp_cm = np.zeros((b, 1, h, w), np.int32)
    np.add.at(exp_cm, (pxpy[:, 0], 0, pxpy[:, 2], pxpy[:, 1]), 1)
    assert eq(cm, exp_cm)
    dense = B.denselize_forward(dev(feats), cm, dev(pxpy))
    exp = np.zeros((b, c, h, w), np.float64)
    for j in range(c):
        np.add.at(exp, (pxpy[:, 0], j, pxpy[:, 2], pxpy[:, 1]),
                  feats[:, j] / exp_cm[pxpy[:, 0], 0, pxpy[:, 2], pxpy[:, 1]])
    assert rel_err(dense, exp) < FP32_TOL
    gd = rng.standard_normal((b, c, h, w)).astype(np.float32)
    gf = B.denselize_backward(dev(gd), cm, dev(pxpy))
    exp_g = gd[pxpy[:, 0], :, pxpy[:, 2], pxpy[:, 1]] / exp_cm[pxpy[:, 0], 0, pxpy[:, 2], pxpy[:, 1]][:, None]
    assert rel_err(gf, exp_g) < FP32_TOL


# ---------------------------------------------------- tensor-core family specific cases
@pytest.mark.parametrize("cin,cout", [(256, 256), (384, 256), (128, 96), (32, 64)])
def test_tc_conv_wide_channels_fp16(ts, cin, cout):
    """tcgen05 path: N > 256 split (dgrad of 384->256), 3 m-tiles in wgrad, SW64 rows (C=32/96)."""
    F = ts.nn.functional
    c = multi_batch_cloud(31, n=900, extent=26, batches=2)
    rng = np.random.default_rng(cin + cout)
    h = lambda a: a.astype(np.float16).astype(np.float32)
    x = h(rng.standard_normal((len(c), cin)))
    w = h(rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin))
    go = h(rng.standard_normal((len(c), cout)))
    nb, ns = R.build_kmap(c, c, 3)
    exp = R.conv_forward(x, w, nb, ns, (len(c), len(c)))
    egi, egw = R.conv_backward(x, w, go, nb, ns)
    xt = dev(x, torch.float16).requires_grad_(True)
    wt = dev(w).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        y = F.conv3d(_sparse(ts, xt, dev(c)), wt, 3)
    assert rel_err(y.feats, exp) < FP16_TOL
    y.feats.backward(dev(go, torch.float16))
    assert rel_err(xt.grad, egi) < FP16_TOL
    assert rel_err(wt.grad, egw) < FP16_TOL


def test_tc_strided_transposed_dense_fp16(ts):
    F = ts.nn.functional
    c = multi_batch_cloud(33, n=1200, extent=28, batches=2)
    rng = np.random.default_rng(9)
    h = lambda a: a.astype(np.float16).astype(np.float32)
    cin, cmid = 64, 96
    x = h(rng.standard_normal((len(c), cin)))
    wd = h(rng.standard_normal((8, cin, cmid)) / np.sqrt(8 * cin))
    wu = h(rng.standard_normal((8, cmid, cin)) / np.sqrt(8 * cmid))
    w1 = h(rng.standard_normal((cin, 128)) / np.sqrt(cin))
    oc = R.spdownsample(c, 2, 2, 1)
    nb, ns = R.build_kmap(c, oc, 2)
    y1 = R.conv_forward(x, wd, nb, ns, (len(c), len(oc)))
    y1h = h(y1)
    y2 = R.conv_forward(y1h, wu, nb, ns, (len(c), len(oc)), transposed=True)
    xt = dev(x, torch.float16).requires_grad_(True)
    wdt, wut, w1t = (dev(a).requires_grad_(True) for a in (wd, wu, w1))
    with torch.autocast("cuda", dtype=torch.float16):
        s0 = _sparse(ts, xt, dev(c))
        s1 = F.conv3d(s0, wdt, 2, stride=2)
        s2 = F.conv3d(s1, wut, 2, stride=2, transposed=True)
        s3 = F.conv3d(s2, w1t, 1)
    assert eq(s1.coords, oc)
    assert rel_err(s1.feats, y1) < FP16_TOL
    assert rel_err(s2.feats, y2) < 2 * FP16_TOL          # two fp16 roundings deep
    y3 = h(s2.feats.detach().float().cpu().numpy()) @ w1
    assert rel_err(s3.feats, y3) < FP16_TOL
    go = h(rng.standard_normal((len(c), 128)))
    s3.feats.backward(dev(go, torch.float16))
    # gradients of the last (dense) layer against numpy on the same fp16 inputs
    s2n = s2.feats.detach().float().cpu().numpy()
    assert rel_err(w1t.grad, s2n.T @ go) < FP16_TOL
    g2 = h(go @ w1.T)
    egi, egw = R.conv_backward(y1h, wu, g2, nb, ns, transposed=True)
    assert rel_err(wut.grad, egw) < 2 * FP16_TOL
    assert torch.isfinite(xt.grad).all() and torch.isfinite(wdt.grad).all()


def test_tc_matches_simt_bitwise_maps_and_close_values(ts):
    """Same inputs through both kernel families (B2S_FORCE_SIMT is read once per process, so
    the SIMT result comes from the fp32 path on the fp16-rounded operands)."""
    F = ts.nn.functional
    c = multi_batch_cloud(35, n=2500, extent=40, batches=3)
    rng = np.random.default_rng(4)
    h = lambda a: a.astype(np.float16).astype(np.float32)
    x, w = h(rng.standard_normal((len(c), 64))), h(rng.standard_normal((27, 64, 64)) / 40)
    y32 = F.conv3d(_sparse(ts, dev(x), dev(c)), dev(w), 3).feats
    with torch.autocast("cuda", dtype=torch.float16):
        y16 = F.conv3d(_sparse(ts, dev(x, torch.float16), dev(c)), dev(w), 3).feats
    assert rel_err(y16, y32) < FP16_TOL


# ----------------------------------------------------------------- fused batch norm
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("c,relu,with_res", [(32, True, False), (96, True, True), (256, False, False),
                                              (64, False, True)])
def test_fused_batch_norm_act(ts, dtype, c, relu, with_res):
    F = ts.nn.functional
    torch.manual_seed(c)
    n = 5000
    tol = 2e-5 if dtype == torch.float32 else 2e-3
    x = (torch.randn(n, c, device="cuda") * 2 + 0.5).to(dtype)
    res = torch.randn(n, c, device="cuda").to(dtype) if with_res else None
    dy = torch.randn(n, c, device="cuda").to(dtype)
    bn_a, bn_b = torch.nn.BatchNorm1d(c).cuda(), torch.nn.BatchNorm1d(c).cuda()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.uniform_(-0.5, 0.5)
    bn_b.load_state_dict(bn_a.state_dict())
    xa, xb = x.clone().requires_grad_(True), x.float().clone().requires_grad_(True)
    ra = res.clone().requires_grad_(True) if with_res else None
    rb = res.float().clone().requires_grad_(True) if with_res else None
    ya = F.batch_norm_act(xa, bn_a, relu=relu, residual=ra)
    yb = bn_b(xb)                                   # fp32 torch reference on the same (rounded) inputs
    if with_res:
        yb = yb + rb
    if relu:
        yb = torch.relu(yb)
    assert ya.dtype == dtype and rel_err(ya, yb) < tol
    ya.backward(dy)
    yb.backward(dy.float())
    assert rel_err(xa.grad, xb.grad) < 5 * tol
    if with_res:
        assert rel_err(ra.grad, rb.grad) < tol
    assert rel_err(bn_a.weight.grad, bn_b.weight.grad) < 5 * tol
    assert rel_err(bn_a.bias.grad, bn_b.bias.grad) < 5 * tol
    assert rel_err(bn_a.running_mean, bn_b.running_mean) < tol
    assert rel_err(bn_a.running_var, bn_b.running_var) < tol
    assert int(bn_a.num_batches_tracked) == 1
    # eval mode falls back to the stock module
    bn_a.eval()
    assert F.batch_norm_act(x, bn_a, relu=relu).shape == x.shape


# ------------------------------------------------------------------- scatter-max (a18)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_scatter_max_matches_torch(dtype):
    from openpcseg_b200.torch_scatter import scatter_max
    torch.manual_seed(0)
    n, c, m = 6000, 24, 500
    src = torch.randn(n, c, device="cuda").to(dtype)
    src[::50] = -3.0                                       # negative maxima and ties
    index = torch.randint(0, m - 7, (n,), device="cuda")   # the last 7 voxels stay empty
    a = src.clone().requires_grad_(True)
    out, arg = scatter_max(a, index, dim=0, dim_size=m)
    ref = torch.full((m, c), float("-inf"), device="cuda", dtype=torch.float32)
    ref.scatter_reduce_(0, index[:, None].expand(-1, c), src.float(), "amax", include_self=True)
    filled = torch.isfinite(ref)
    assert torch.equal(out.float()[filled], ref[filled]) and bool((out[~filled] == 0).all())
    assert bool((arg[~filled] == n).all())
    # arg points at a row of the right voxel holding the maximum (the smallest such row)
    rows = arg[filled]
    cols = torch.arange(c, device="cuda").expand(m, c)[filled]
    assert torch.equal(src[rows, cols].float(), ref[filled])
    g = torch.randn(m, c, device="cuda").to(dtype)
    out.backward(g)
    exp = torch.zeros(n + 1, c, device="cuda", dtype=dtype)
    exp.scatter_(0, arg, g)
    assert torch.equal(a.grad, exp[:n])


# ------------------------------------ config-4 / config-3 style chains (parity-test cases)
def test_cylinder_style_chain_fp32(ts):
    """Cylinder3D's conv vocabulary in one chain: asymmetric (1,3,3) and (3,1,3) submanifold convs,
    a k3 pooling conv with stride (2,2,1) (slow-path downsample, more outputs than a snap would give)
    and its transposed conv back (cylinder_ts.py:204-214, 292-301), forward and backward vs the oracle."""
    F = ts.nn.functional
    c = multi_batch_cloud(41, n=1500, extent=28, batches=2)
    rng = np.random.default_rng(41)
    ci, cm = 8, 16
    x = rng.standard_normal((len(c), ci)).astype(np.float32)
    w1 = (rng.standard_normal((9, ci, cm)) / 9).astype(np.float32)
    w2 = (rng.standard_normal((9, cm, cm)) / 12).astype(np.float32)
    wp = (rng.standard_normal((27, cm, cm)) / 20).astype(np.float32)
    wt = (rng.standard_normal((27, cm, ci)) / 20).astype(np.float32)
    nb1, ns1 = R.build_kmap(c, c, (1, 3, 3))
    nb2, ns2 = R.build_kmap(c, c, (3, 1, 3))
    oc = R.spdownsample(c, (2, 2, 1), 3, 1)
    nbp, nsp = R.build_kmap(c, oc, 3)
    n, m = len(c), len(oc)
    y1 = R.conv_forward(x, w1, nb1, ns1, (n, n))
    y2 = R.conv_forward(y1, w2, nb2, ns2, (n, n))
    y3 = R.conv_forward(y2, wp, nbp, nsp, (n, m))
    y4 = R.conv_forward(y3, wt, nbp, nsp, (n, m), transposed=True)
    go = rng.standard_normal(y4.shape).astype(np.float32)
    g3, gwt = R.conv_backward(y3, wt, go, nbp, nsp, transposed=True)
    g2, gwp = R.conv_backward(y2, wp, g3, nbp, nsp)
    g1, gw2 = R.conv_backward(y1, w2, g2, nb2, ns2)
    g0, gw1 = R.conv_backward(x, w1, g1, nb1, ns1)
    xt = dev(x).requires_grad_(True)
    ws = [dev(a).requires_grad_(True) for a in (w1, w2, wp, wt)]
    s0 = _sparse(ts, xt, dev(c))
    s1 = F.conv3d(s0, ws[0], (1, 3, 3))
    s2 = F.conv3d(s1, ws[1], (3, 1, 3))
    s3 = F.conv3d(s2, ws[2], 3, stride=(2, 2, 1))
    s4 = F.conv3d(s3, ws[3], 3, stride=(2, 2, 1), transposed=True)
    assert eq(s3.coords, oc) and s3.stride == (2, 2, 1) and eq(s4.coords, c)
    assert rel_err(s3.feats, y3) < FP32_TOL and rel_err(s4.feats, y4) < FP32_TOL
    s4.feats.backward(dev(go))
    assert rel_err(xt.grad, g0) < 2 * FP32_TOL
    for got, exp in zip(ws, (gw1, gw2, gwp, gwt)):
        assert rel_err(got.grad, exp) < 2 * FP32_TOL


def test_spvcnn_style_point_voxel_fusion(ts):
    """SPVCNN's extra hot-path calls: point_to_voxel (scatter-mean of point features onto a strided
    level) feeding a conv, and voxel_to_point of the result, summed with a point branch
    (spvcnn.py:411-433), forward + backward vs the oracle."""
    from openpcseg_b200.segmentors import initial_voxelize, point_to_voxel, voxel_to_point
    F = ts.nn.functional
    rng = np.random.default_rng(5)
    n, cpt = 4000, 16
    pts = np.concatenate([rng.uniform(0, 30, (n, 3)), rng.integers(0, 2, (n, 1))], 1).astype(np.float32)
    pf = rng.standard_normal((n, cpt)).astype(np.float32)
    wd = (rng.standard_normal((8, cpt, cpt)) / 8).astype(np.float32)
    wc = (rng.standard_normal((27, cpt, cpt)) / 20).astype(np.float32)
    # oracle
    vc, vf, idx, cnt, nfc = R.initial_voxelize(pts, pf, 1.0, 1.0)
    oc = R.spdownsample(vc, 2, 2, 1)
    nbd, nsd = R.build_kmap(vc, oc, 2)
    x1 = R.conv_forward(vf, wd, nbd, nsd, (len(vc), len(oc)))
    pi, pc = R.point_to_voxel_map(nfc, oc, 2)
    fused = x1 + R.spvoxelize_forward(pf, pi, pc)
    nbc, nsc = R.build_kmap(oc, oc, 3, 2)
    x2 = R.conv_forward(fused, wc, nbc, nsc, (len(oc), len(oc)))
    i8, w8 = R.trilinear_map(nfc, oc, 2)
    out = R.spdevoxelize_forward(x2, i8, w8) + pf
    # CUDA path
    pft = dev(pf).requires_grad_(True)
    z = ts.PointTensor(pft, dev(pts))
    v0 = initial_voxelize(z, 1.0, 1.0)
    assert eq(v0.C, vc)
    v1 = F.conv3d(v0, dev(wd), 2, stride=2)
    zf = ts.PointTensor(pft, z.C, idx_query=z.idx_query, weights=z.weights)
    zf.additional_features = z.additional_features
    v1 = v1 + point_to_voxel(v1, zf)
    v2 = F.conv3d(v1, dev(wc), 3)
    res = voxel_to_point(v2, zf).F + pft
    assert rel_err(res, out) < 2 * FP32_TOL
    res.square().sum().backward()
    # d/d(point feats) through voxelize (twice), both convs and devoxelize, against the oracle chain
    g = 2 * out
    gx2 = R.spdevoxelize_backward(g, i8, w8, len(oc))
    gfused, _ = R.conv_backward(fused, wc, gx2, nbc, nsc)
    gvf, _ = R.conv_backward(vf, wd, gfused, nbd, nsd)
    gp = g + R.spvoxelize_backward(gfused, pi, pc, n) + R.spvoxelize_backward(gvf, idx, cnt, n)
    assert rel_err(pft.grad, gp) < 5 * FP32_TOL
