"""Freeze the reference's LR-schedule multipliers and optimizer wiring (run in the build container, where
/root/reference exists):  python tests/golden/make_golden_optim.py  ->  tests/golden/optim.npz"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import pcseg.optim as O                                              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    steps = np.arange(0, 3600, 7)
    warm, total = 100, 3600
    for name in ("linear_warmup_with_cosdecay", "cos_warmup_with_cosdecay"):
        out[name] = np.array([getattr(O, name)(int(s), warm, total) for s in steps])
    dsteps, dscales = [1200, 2400], [0.1, 0.5]
    out["linear_warmup_with_stepdecay"] = np.array(
        [O.linear_warmup_with_stepdecay(int(s), warm, total, dsteps, dscales) for s in steps])
    out["coswarmup_with_stepdecay"] = np.array(
        [O.coswarmup_with_stepdecay(int(s), warm, total, dsteps, dscales) for s in steps])
    out["steps"], out["warm_total"], out["decay_steps"], out["decay_scales"] = steps, np.array([warm, total]), \
        np.array(dsteps), np.array(dscales)
    # LambdaLR wiring: lr trajectory of the reference's build_scheduler on a toy model
    cfg = types.SimpleNamespace(OPTIMIZER="sgd", LR=0.24, WEIGHT_DECAY=1e-4, MOMENTUM=0.9,
                                SCHEDULER="linear_warmup_with_cosdecay", WARMUP_EPOCH=1)
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 3)
    opt = O.build_optimizer(model, cfg)
    sched = O.build_scheduler(opt, total_iters_each_epoch=10, total_epochs=5, optim_cfg=cfg)
    lrs = []
    for _ in range(50):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    out["lambda_lr"] = np.array(lrs)
    g = opt.param_groups[0]
    out["sgd_group"] = np.array([g["momentum"], g["weight_decay"], float(g["nesterov"]), g["dampening"]])
    np.savez(os.path.join(HERE, "optim.npz"), **out)
    print("wrote optim.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
