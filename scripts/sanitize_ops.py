#!/usr/bin/env python
"""One small invocation of every libb2s kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_ops.py
    compute-sanitizer --tool racecheck python scripts/sanitize_ops.py
    compute-sanitizer --tool synccheck python scripts/sanitize_ops.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import openpcseg_b200.torchsparse as ts
    from openpcseg_b200 import backend as B
    from openpcseg_b200.synthetic import make_batch
    from openpcseg_b200.segmentors.point_voxel import initial_voxelize, voxel_to_point
    F = ts.nn.functional
    b = make_batch([1, 2], n_azimuth=40)
    c = torch.from_numpy(b["coords"]).cuda()
    n = c.shape[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    for cin, cout in ((32, 32), (96, 96), (64, 256)):
        x = torch.randn(n, cin, device="cuda", generator=g).half().requires_grad_(True)
        w = (torch.randn(27, cin, cout, device="cuda", generator=g) / 30).requires_grad_(True)
        st = ts.SparseTensor(x, c, 1)
        st.cmaps[st.stride] = c
        with torch.autocast("cuda", dtype=torch.float16):
            sums = torch.zeros(2, cout, dtype=torch.float64, device="cuda")
            y = F.conv3d(st, w, 3, bn_sums=sums)
            y2 = F.conv3d(y, torch.randn(8, cout, 32, device="cuda", generator=g) / 10, 2, stride=2)
            y3 = F.conv3d(y2, torch.randn(8, 32, cout, device="cuda", generator=g) / 10, 2, stride=2, transposed=True)
        bn = torch.nn.BatchNorm1d(cout).cuda().train()
        z = F.batch_norm_act(y.feats, bn, relu=True, residual=y3.feats)
        z.float().sum().backward()
    xf = torch.randn(n, 16, device="cuda", generator=g).requires_grad_(True)
    st = ts.SparseTensor(xf, c, 1)
    st.cmaps[st.stride] = c
    F.conv3d(st, torch.randn(27, 16, 16, device="cuda", generator=g).requires_grad_(True), 3).feats.sum().backward()
    pt = ts.PointTensor(torch.from_numpy(b["feats"]).cuda(), c.float())
    x0 = initial_voxelize(pt, 0.05, 0.05)
    x1 = F.conv3d(x0, torch.randn(8, 4, 32, device="cuda", generator=g).requires_grad_(True), 2, stride=2)
    z1 = voxel_to_point(x1, pt)
    z1.F.sum().backward()
    B.scatter_max(torch.randn(n, 32, device="cuda", generator=g), torch.randint(0, n // 4, (n,), device="cuda"), n // 4)
    torch.cuda.synchronize()
    print("sanitize_ops: done,", n, "voxels")


if __name__ == "__main__":
    main()
