"""Synchronised fused batch norm (_SyncBatchNormAct: one all-reduce of fp64 sums per direction) on two ranks
== the single-process fused batch norm over the concatenated rows.  Both ranks share cuda:0 and talk through
gloo (NCCL refuses two ranks on one device), so the test runs on the 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(c):
    g = torch.Generator().manual_seed(c)
    n = 3000
    x = torch.randn(n, c, generator=g).half()
    res = torch.randn(n, c, generator=g).half()
    dy = torch.randn(n, c, generator=g).half()
    return x, res, dy


def _worker(rank, world, port, c, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from openpcseg_b200.torchsparse.nn.functional import batch_norm_act
        torch.cuda.set_device(0)
        x, res, dy = _data(c)
        cut = 1100                                            # unequal shards: the global count matters
        sl = slice(0, cut) if rank == 0 else slice(cut, None)
        bn = torch.nn.SyncBatchNorm(c).cuda().train()
        with torch.no_grad():
            bn.weight.copy_(torch.linspace(0.5, 1.5, c))
            bn.bias.copy_(torch.linspace(-0.2, 0.2, c))
        xs = x[sl].cuda().requires_grad_(True)
        rs = res[sl].cuda().requires_grad_(True)
        y = batch_norm_act(xs, bn, relu=True, residual=rs)
        y.backward(dy[sl].cuda())
        # numpy payloads are pickled by value (torch tensors travel as shared-memory handles that die with the worker)
        out.put((rank,) + tuple(t.detach().float().cpu().numpy() for t in
                                (y, xs.grad, rs.grad, bn.weight.grad, bn.bias.grad, bn.running_mean, bn.running_var)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("c", [32, 96])
def test_sync_batchnorm_two_ranks_equals_full_batch(c):
    from openpcseg_b200.torchsparse.nn.functional import batch_norm_act
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, c, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([out.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    got = [(g[0],) + tuple(torch.from_numpy(a) for a in g[1:]) for g in got]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, res, dy = _data(c)
    bn = torch.nn.BatchNorm1d(c).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(torch.linspace(0.5, 1.5, c))
        bn.bias.copy_(torch.linspace(-0.2, 0.2, c))
    xf, rf = x.cuda().requires_grad_(True), res.cuda().requires_grad_(True)
    y = batch_norm_act(xf, bn, relu=True, residual=rf)
    y.backward(dy.cuda())
    cat = lambda i: torch.cat([got[0][i], got[1][i]])
    assert torch.allclose(cat(1), y.detach().float().cpu(), atol=2e-3, rtol=2e-3)
    assert torch.allclose(cat(2), xf.grad.float().cpu(), atol=2e-3, rtol=2e-3)
    assert torch.allclose(cat(3), rf.grad.float().cpu(), atol=2e-3, rtol=2e-3)
    # local weight / bias gradients add up to the full-batch ones (DDP would average them)
    assert torch.allclose(got[0][4] + got[1][4], bn.weight.grad.cpu(), atol=2e-2, rtol=2e-3)
    assert torch.allclose(got[0][5] + got[1][5], bn.bias.grad.cpu(), atol=2e-2, rtol=2e-3)
    for r in range(2):                                       # identical running statistics on every rank
        assert torch.allclose(got[r][6], bn.running_mean.cpu(), atol=1e-5)
        assert torch.allclose(got[r][7], bn.running_var.cpu(), atol=1e-5)
