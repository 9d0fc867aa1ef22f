"""world_size-2 gloo tests (CPU) of the data-parallel plumbing used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from openpcseg_b200 import dist_utils as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert D.env_rank_world() == (rank, world, rank)
        # timing: max over ranks
        ms = D.max_over_ranks(10.0 + 5.0 * rank, torch.device("cpu"))
        # DDP gradient all-reduce == mean of per-rank gradients
        torch.manual_seed(0)
        model = torch.nn.Linear(4, 3)
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        x = torch.full((2, 4), float(rank + 1))
        ddp(x).sum().backward()
        grad = model.weight.grad.clone()
        seeds = D.scan_seeds(rank, 1, 3)
        gathered = [None] * world
        dist.all_gather_object(gathered, seeds)
        if rank == 0:
            out.put({"ms": ms, "grad": grad.tolist(), "seeds": gathered})
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ms"] == 15.0
    # d(sum(Wx+b))/dW = sum over batch of x = 2*(rank+1) per column; DDP averages ranks -> 3
    assert torch.allclose(torch.tensor(res["grad"]), torch.full((3, 4), 3.0))
    flat = [s for ss in res["seeds"] for s in ss]
    assert len(set(flat)) == len(flat) == 6


def test_whole_job_rate():
    from openpcseg_b200.dist_utils import whole_job_rate
    assert whole_job_rate(8, 4, 2000.0) == 16.0
