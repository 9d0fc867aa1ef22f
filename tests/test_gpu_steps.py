"""Step table of the tensor-core convolution (b2s_tile_steps, consumed by conv_tc4.cu): layout against a numpy
restatement of include/b2s.h, and the batch-norm statistics the conv epilogue accumulates."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _expect(nbr, perm, tr):
    k, n = nbr.shape
    tiles = max((n + tr - 1) // tr, 1)
    words = (k + 31) // 32
    mask = np.zeros((tiles, words), np.uint32)
    start = np.zeros(tiles + 1, np.int64)
    rows = []
    for t in range(tiles):
        cols = [(perm[j] if perm is not None else j) if j < n else -1 for j in range(t * tr, (t + 1) * tr)]
        sub = np.stack([np.where(np.array(cols) >= 0, nbr[kk][np.maximum(cols, 0)], -1) for kk in range(k)])
        active = [kk for kk in range(k) if (sub[kk] >= 0).any()]
        for kk in active:
            mask[t, kk // 32] |= np.uint32(1 << (kk % 32))
            slot = np.full(tr, -1, np.int64)
            for row in range(tr):
                w, i, q = row // 32, (row % 32) // 4, row % 4
                slot[w * 32 + q * 8 + i] = sub[kk][row]
            rows.append(slot)
        start[t + 1] = start[t] + len(active)
    return mask, start, (np.stack(rows) if rows else np.zeros((0, tr), np.int64))


@pytest.mark.parametrize("tr", [128, 256])
@pytest.mark.parametrize("k,n,use_perm", [(27, 1000, True), (27, 300, False), (8, 2049, False), (40, 700, True)])
def test_tile_steps_layout(tr, k, n, use_perm):
    from openpcseg_b200 import backend as B
    rng = np.random.default_rng(k * n + tr)
    nbr = rng.integers(0, n, size=(k, n)).astype(np.int32)
    keep = rng.random((k, n)) < 0.15
    keep[:, : n // 3] &= (np.arange(k) % 4 == 0)[:, None]          # tiles with few active offsets
    keep[k // 2, n // 2:] = False                                    # an offset no late tile needs
    nbr = np.where(keep, nbr, -1).astype(np.int32)
    perm = rng.permutation(n).astype(np.int32) if use_perm else None
    bits = None
    if use_perm and k <= 32:
        bits = np.zeros(n, np.uint32)
        for kk in range(k):
            bits |= np.where(nbr[kk] >= 0, np.uint32(1 << kk), np.uint32(0))
    mask, start, rows, tr_out = B.tile_steps(torch.from_numpy(nbr).cuda(),
                                             torch.from_numpy(perm).cuda() if perm is not None else None,
                                             torch.from_numpy(bits.view(np.int32)).cuda() if bits is not None else None, tr)
    e_mask, e_start, e_rows = _expect(nbr, perm, tr)
    assert tr_out == tr
    assert np.array_equal(mask.cpu().numpy().view(np.uint32), e_mask)
    assert np.array_equal(start.cpu().numpy(), e_start)
    got = rows.cpu().numpy()[: e_rows.size].reshape(-1, tr)
    assert np.array_equal(got, e_rows)


@pytest.mark.parametrize("c_in,c_out", [(32, 32), (64, 96), (96, 96), (128, 256)])
def test_conv_epilogue_batchnorm_sums(c_in, c_out):
    """bn_sums of b2s_conv_gather_gemm_steps == per-channel sum / sum of squares of the fp16 rows it wrote."""
    import openpcseg_b200.torchsparse as ts
    from openpcseg_b200 import backend as B
    from openpcseg_b200.synthetic import make_batch
    F = ts.nn.functional
    c = torch.from_numpy(make_batch([5], n_azimuth=500)["coords"]).cuda()
    n = c.shape[0]
    km = F.build_kernel_map(c, c, 3, (1, 1, 1))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, c_in, device="cuda", generator=g).half()
    w = (torch.randn(27, c_in, c_out, device="cuda", generator=g) / (27 * c_in) ** 0.5).half()
    n_rows, kw = km.gather_args("out", x, c_in, c_out)
    assert "steps" in kw
    sums = torch.zeros(2, c_out, dtype=torch.float64, device="cuda")
    y = B.conv_gather_gemm(x, w, n_rows=n_rows, transpose_w=False, bn_sums=sums, **kw)
    y2 = B.conv_gather_gemm(x, w, n_rows=n_rows, transpose_w=False, **kw)
    assert torch.equal(y, y2)
    yd = y.double()
    assert torch.allclose(sums[0], yd.sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(sums[1], (yd * yd).sum(0), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("use_perm", [False, True])
def test_chunked_pair_list_and_segmented_wgrad(use_perm):
    """b2s_kmap_pairs_chunked emits every pair exactly once in (row range, offset, row) order, and the weight
    gradient over that segmented list equals the one over the reference-order list."""
    import openpcseg_b200.torchsparse as ts
    from openpcseg_b200 import backend as B
    from openpcseg_b200.synthetic import make_batch
    F = ts.nn.functional
    c = torch.from_numpy(make_batch([2, 3], n_azimuth=300)["coords"]).cuda()
    n = c.shape[0]
    km = F.build_kernel_map(c, c, 3, (1, 1, 1))
    k, n_chunks = km.kvol, 5
    perm = F._tile_order(c) if use_perm else None
    pairs, seg, total = B.kmap_pairs_chunked(km.nbr_out, perm, n_chunks)
    m = int(total.item())
    ref_pairs, ref_total = km.pairs()
    assert m == int(ref_total.item()) == int(seg.sum().item())
    got = pairs[:m].cpu().numpy()
    nbr = km.nbr_out.cpu().numpy()
    chunk_rows = -(-(-(-n // n_chunks)) // 32) * 32
    rank = np.empty(n, np.int64)
    rank[(perm.cpu().numpy() if perm is not None else np.arange(n))] = np.arange(n)
    pos = 0
    segs = seg.cpu().numpy().reshape(n_chunks, k)
    for ch in range(n_chunks):
        for kk in range(k):
            blk = got[pos:pos + segs[ch, kk]]
            pos += segs[ch, kk]
            assert np.array_equal(nbr[kk][blk[:, 1]], blk[:, 0])               # a pair of offset kk
            r = rank[blk[:, 1]]
            assert ((r // chunk_rows) == ch).all() and (np.diff(r) > 0).all()   # in its range, launch order
    a = sorted(map(tuple, got.tolist()))
    b = sorted(map(tuple, ref_pairs[:m].cpu().numpy().tolist()))
    assert a == b
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, 64, device="cuda", generator=g).half()
    gy = torch.randn(n, 96, device="cuda", generator=g).half()
    w_ref = B.conv_wgrad(x, gy, k, ref_pairs, km.nbsizes32, False)
    w_seg = B.conv_wgrad(x, gy, k, pairs, seg, False)
    assert float((w_ref - w_seg).abs().max()) <= 1e-4 * float(w_ref.abs().max())


def test_weight_bank_refreshes_all_parameters_in_one_launch():
    """The fp16 operand copies of the conv parameters (functional._WeightBank, b2s_weights_refresh): bit-exact
    round-to-nearest casts in both layouts at first sight, refreshed for EVERY registered parameter when one is
    looked up after an in-place update, with odd shapes (tile edges), a 2-D (1x1x1) weight and a dropped one."""
    import gc
    from openpcseg_b200.torchsparse.nn import functional as F
    g = torch.Generator().manual_seed(3)
    shapes = [(27, 96, 96), (8, 32, 64), (27, 40, 72), (1, 33, 17), (64, 128)]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]

    def check(p):
        cast = F._cast_weight(p, torch.float16)
        km = F._kmajor_weight(p, cast)
        p3 = p.detach() if p.ndim == 3 else p.detach().unsqueeze(0)
        assert cast.dtype == torch.float16 and torch.equal(cast, p.detach().half())
        assert km.shape == (p3.shape[0], p3.shape[2], p3.shape[1]) and torch.equal(km, p3.half().transpose(1, 2))
        return cast, km

    bufs = [check(p) for p in params]
    bank = F._WEIGHT_BANKS[params[0].device]
    with torch.no_grad():
        for p in params:
            p.mul_(0.5).add_(0.123)                                   # optimizer-style in-place update
    first = check(params[2])                                          # stale -> one launch refreshes all
    n0 = len(bank.entries)                                            # (dead entries of earlier tests are gone now)
    assert n0 >= len(params)
    assert all(e.key == (e.ref()._version, e.ref().data_ptr()) for e in bank.entries if e.ref() is not None)
    for p, (cast, km) in zip(params, bufs):
        c2, k2 = check(p)
        assert c2.data_ptr() == cast.data_ptr() and k2.data_ptr() == km.data_ptr()      # persistent buffers
    assert first[0].data_ptr() == bufs[2][0].data_ptr()
    dead = params.pop(1)
    del dead, bufs
    gc.collect()
    with torch.no_grad():
        params[0].add_(1.0)
    check(params[0])
    assert len(bank.entries) == n0 - 1                               # the dropped parameter left the table
    for p in params:
        check(p)
    # a non-parameter (e.g. the zero-padded copy of a narrow layer) takes the per-call path
    t = torch.randn(27, 32, 32, generator=g).cuda()
    assert F._banked(t, torch.float16) is None
    assert torch.equal(F._cast_weight(t, torch.float16), t.half())


def test_zero_sums_slices_are_fresh_and_disjoint():
    from openpcseg_b200.torchsparse.nn import functional as F
    dev = torch.device("cuda", 0)
    seen = []
    for i in range(700):                                              # crosses several chunks
        s = F.zero_sums(32 + 32 * (i % 8), dev)
        assert s.shape == (2, 32 + 32 * (i % 8)) and s.dtype == torch.float64 and s.is_contiguous()
        assert float(s.abs().sum()) == 0.0
        s.fill_(float(i + 1))
        seen.append(s)
    for i, s in enumerate(seen):
        assert bool((s == float(i + 1)).all())
