"""Size-independent properties of the path, checked at BASELINE's full scan size on the CUDA path
(`-m gpu`) and - so that the checkers themselves are pinned - on the oracle at a small size (CPU).

Kernel maps: pair convention in = out + offset[k], pairs sorted (k, out), centre offset = identity,
mirror symmetry of submanifold maps, k2s2 partitions the inputs.  Convolution: linearity, and the three
adjoint identities <conv(x), g> = <x, dgrad(g)> = <W, wgrad(x, g)>.  Hash table: query(build(keys)) = id.
"""
import functools

import numpy as np
import pytest
import torch

from oracle import ref_ops as R


# ------------------------------------------------------------------ device-agnostic checkers
def check_submanifold_map(coords, nbmaps, nbsizes, offsets, nbr_out=None):
    """coords int [N,4]; nbmaps int64 [M,2] = (in, out); nbsizes [27]; offsets [27,3]."""
    n, kvol = coords.shape[0], offsets.shape[0]
    nbsizes = nbsizes.long()
    assert int(nbsizes.sum()) == nbmaps.shape[0]
    k_of = torch.repeat_interleave(torch.arange(kvol, device=coords.device), nbsizes)
    cin, cout = coords[nbmaps[:, 0]].long(), coords[nbmaps[:, 1]].long()
    assert torch.equal(cin[:, :3], cout[:, :3] + offsets.long()[k_of])           # in = out + offset[k]
    assert torch.equal(cin[:, 3], cout[:, 3])                                    # never across scans
    same_k = k_of[1:] == k_of[:-1]
    assert bool(((nbmaps[1:, 1] > nbmaps[:-1, 1]) | ~same_k).all())              # out ascending inside an offset
    centre = kvol // 2
    lo = int(nbsizes[:centre].sum())
    assert int(nbsizes[centre]) == n
    assert torch.equal(nbmaps[lo:lo + n, 0], torch.arange(n, device=coords.device))
    assert torch.equal(nbmaps[lo:lo + n, 1], torch.arange(n, device=coords.device))
    assert torch.equal(nbsizes, nbsizes.flip(0))                                 # offset k <-> offset K-1-k
    if nbr_out is not None:                                                      # (i, o) at k  <=>  (o, i) at K-1-k
        assert torch.equal(nbr_out[kvol - 1 - k_of, nbmaps[:, 0]].long(), nbmaps[:, 1])
        assert int((nbr_out >= 0).sum()) == nbmaps.shape[0]


def check_k2s2_map(coords, out_coords, nbmaps, nbsizes):
    n = coords.shape[0]
    assert bool((out_coords[:, :3] % 2 == 0).all())
    key = (out_coords[:, 3].long() << 54) | (out_coords[:, 0].long() << 36) | (out_coords[:, 1].long() << 18) | \
        out_coords[:, 2].long()
    assert bool((key[1:] > key[:-1]).all())                                      # unique, sorted (b, x, y, z)
    assert int(nbsizes.sum()) == n == nbmaps.shape[0]                            # every input in exactly one cell
    assert torch.equal(torch.sort(nbmaps[:, 0]).values, torch.arange(n, device=coords.device))
    cin, cout = coords[nbmaps[:, 0]].long(), out_coords[nbmaps[:, 1]].long()
    assert torch.equal(cin[:, :3] // 2 * 2, cout[:, :3]) and torch.equal(cin[:, 3], cout[:, 3])


def dot64(a, b):
    return float((a.double() * b.double()).sum())


def check_linearity_and_adjoints(conv, x1, x2, w, go, tol):
    """conv(x, w) -> (y, grad_x, grad_w) for grad_out = go."""
    y1, gx1, gw1 = conv(x1, w)
    y2, _, _ = conv(x2, w)
    y12, _, _ = conv(2.0 * x1 - 3.0 * x2, w)
    scale = float(y12.abs().max())
    assert float((y12 - (2.0 * y1 - 3.0 * y2)).abs().max()) <= tol * scale
    ref = dot64(y1, go)
    assert abs(dot64(x1, gx1) - ref) <= tol * abs(ref) + tol * float(y1.abs().max()) * float(go.abs().max())
    assert abs(dot64(w, gw1) - ref) <= tol * abs(ref) + tol * float(y1.abs().max()) * float(go.abs().max())


# --------------------------------------------------------------------------- CPU: pin the checkers
def _cloud(seed, n=3000):
    rng = np.random.default_rng(seed)
    parts = []
    for b in range(2):
        c = np.unique(rng.integers(0, 24, size=(n, 3)), axis=0)
        parts.append(np.concatenate([c, np.full((len(c), 1), b)], 1))
    return np.concatenate(parts).astype(np.int32)


def test_checkers_accept_the_oracle():
    c = _cloud(0)
    nb, ns = R.build_kmap(c, c, 3)
    offs = R.get_kernel_offsets(3, 1)
    check_submanifold_map(torch.from_numpy(c), torch.from_numpy(nb).long(), torch.from_numpy(ns),
                          torch.from_numpy(np.asarray(offs)))
    oc = R.spdownsample(c, 2, 2, 1)
    nb2, ns2 = R.build_kmap(c, oc, 2)
    check_k2s2_map(torch.from_numpy(c), torch.from_numpy(oc), torch.from_numpy(nb2).long(), torch.from_numpy(ns2))
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((27, 8, 6)) / 10).astype(np.float32)
    go = rng.standard_normal((len(c), 6)).astype(np.float32)

    def conv(x, wt):
        y = R.conv_forward(x.numpy(), wt.numpy(), nb, ns, (len(c), len(c)))
        gx, gw = R.conv_backward(x.numpy(), wt.numpy(), go, nb, ns)
        return torch.from_numpy(y), torch.from_numpy(gx), torch.from_numpy(gw)

    x1, x2 = (torch.from_numpy(rng.standard_normal((len(c), 8)).astype(np.float32)) for _ in range(2))
    check_linearity_and_adjoints(conv, x1, x2, torch.from_numpy(w), torch.from_numpy(go), 1e-5)


def test_checkers_reject_a_broken_map():
    c = _cloud(2)
    nb, ns = R.build_kmap(c, c, 3)
    bad = nb.copy()
    bad[5, 0] = (bad[5, 0] + 1) % len(c)
    with pytest.raises(AssertionError):
        check_submanifold_map(torch.from_numpy(c), torch.from_numpy(bad).long(), torch.from_numpy(ns),
                              torch.from_numpy(np.asarray(R.get_kernel_offsets(3, 1))))


# --------------------------------------------------------------------------- GPU: full scan size
@functools.lru_cache(maxsize=1)
def _full_batch_host():
    from openpcseg_b200.synthetic import make_batch
    return make_batch([0, 1])["coords"]                                   # 2 scans, ~190 k voxels


def _full_batch():
    return torch.from_numpy(_full_batch_host()).cuda()


@pytest.mark.gpu
def test_full_size_kernel_maps_and_hash_table():
    import openpcseg_b200.torchsparse as ts
    F = ts.nn.functional
    c = _full_batch()
    assert c.shape[0] > 150_000
    km = F.build_kernel_map(c, c, 3, (1, 1, 1))
    nbmaps, nbsizes, sizes = km
    assert sizes == (c.shape[0], c.shape[0])
    check_submanifold_map(c, nbmaps, nbsizes, ts.nn.utils.get_kernel_offsets(3, 1, 1, "cuda"), km.nbr_out)
    oc = F.spdownsample(c, 2, 2, 1)
    nb2, ns2, sz2 = F.build_kernel_map(c, oc, 2, (1, 1, 1))
    assert sz2 == (c.shape[0], oc.shape[0])
    check_k2s2_map(c, oc, nb2, ns2)
    h = F.sphash(c)
    assert torch.equal(F.sphashquery(h, h), torch.arange(c.shape[0], device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout", [(16, 16), (96, 96)])
def test_full_size_conv_linearity_and_adjoints_fp32(cin, cout):
    import openpcseg_b200.torchsparse as ts
    F = ts.nn.functional
    c = _full_batch()
    g = torch.Generator(device="cuda").manual_seed(cin)
    x1 = torch.randn(c.shape[0], cin, device="cuda", generator=g)
    x2 = torch.randn(c.shape[0], cin, device="cuda", generator=g)
    w = torch.randn(27, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5
    go = torch.randn(c.shape[0], cout, device="cuda", generator=g)

    def conv(x, wt):
        xt, wl = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
        st = ts.SparseTensor(xt, c, 1)
        st.cmaps[st.stride] = st.coords
        y = F.conv3d(st, wl, 3).feats
        y.backward(go)
        return y.detach(), xt.grad, wl.grad

    check_linearity_and_adjoints(conv, x1, x2, w, go, 1e-4)


@pytest.mark.gpu
def test_full_size_fp16_tensor_core_path_against_fp32_path():
    import openpcseg_b200.torchsparse as ts
    F = ts.nn.functional
    c = _full_batch()
    g = torch.Generator(device="cuda").manual_seed(7)
    h = lambda t: t.half().float()
    x = h(torch.randn(c.shape[0], 96, device="cuda", generator=g))
    w = h(torch.randn(27, 96, 96, device="cuda", generator=g) / 50)
    go = h(torch.randn(c.shape[0], 96, device="cuda", generator=g))

    def run(half):
        xt = (x.half() if half else x.clone()).requires_grad_(True)
        wl = w.clone().requires_grad_(True)
        st = ts.SparseTensor(xt, c, 1)
        st.cmaps[st.stride] = st.coords
        with torch.autocast("cuda", dtype=torch.float16, enabled=half):
            y = F.conv3d(st, wl, 3).feats
        y.backward(go.half() if half else go)
        return y.float(), xt.grad.float(), wl.grad.float()

    ref, out = run(False), run(True)
    for a, b in zip(out, ref):
        assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout", [(96, 96), (256, 256), (32, 32), (128, 96)])
def test_full_size_fp16_conv_against_the_reference_cpu_backend(cin, cout):
    """BASELINE-size (2 scans, ~190 k voxels = ~1500 row tiles per launch: the persistent multi-tile loop, ring
    wrap-around across tiles, TMEM accumulator turn-taking and the row_perm epilogue) fp16 tensor-core forward,
    input gradient and weight gradient against an INDEPENDENT answer: the reference's own compiled CPU backend
    (oracle/_ref: convolution_forward_cpu / convolution_backward_cpu, fp32) on the same fp16-representable
    operands and the same reference-format pair list.  Bar: 1e-3 of the result's max (north_star, fp16)."""
    import openpcseg_b200.torchsparse as ts
    from oracle import build_ref
    F = ts.nn.functional
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref (reference CPU backend) is not built")
    torch.set_num_threads(8)
    c = _full_batch()
    n = c.shape[0]
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    h = lambda t: t.half().float()
    x = h(torch.randn(n, cin, device="cuda", generator=g))
    w = h(torch.randn(27, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5 * 3)
    go = h(torch.randn(n, cout, device="cuda", generator=g))
    xt, wl = x.half().requires_grad_(True), w.clone().requires_grad_(True)
    st = ts.SparseTensor(xt, c, 1)
    st.cmaps[st.stride] = st.coords
    with torch.autocast("cuda", dtype=torch.float16):
        y = F.conv3d(st, wl, 3)
    y.feats.backward(go.half())
    nbmaps, nbsizes, _ = y.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    nb, ns = nbmaps.int().cpu().contiguous(), nbsizes.int().cpu().contiguous()
    xc, wc, gc = x.cpu(), w.cpu(), go.cpu()
    out = torch.zeros(n, cout)
    ref.convolution_forward_cpu(xc, out, wc, nb, ns, False)
    gi, gw = torch.zeros_like(xc), torch.zeros_like(wc)
    ref.convolution_backward_cpu(xc, gi, gc, wc, gw, nb, ns, False)
    for name, a, b in (("fwd", y.feats, out), ("dgrad", xt.grad, gi), ("wgrad", wl.grad, gw)):
        err = float((a.float().cpu() - b).abs().max() / b.abs().max())
        assert err <= 1e-3, (name, cin, cout, err)
