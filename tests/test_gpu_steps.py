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
