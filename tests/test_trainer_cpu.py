"""Training-step driver against values frozen from the reference's pcseg/optim (tests/golden/optim.npz,
generator: tests/golden/make_golden_optim.py)."""
import io
import os

import numpy as np
import torch

from openpcseg_b200.trainer import OptimConfig, TrainStep, build_optimizer, build_scheduler, schedule_multiplier

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "optim.npz"))


def test_schedule_multipliers_match_reference():
    warm, total = (int(v) for v in G["warm_total"])
    for name in ("linear_warmup_with_cosdecay", "cos_warmup_with_cosdecay", "linear_warmup_with_stepdecay",
                 "coswarmup_with_stepdecay"):
        got = [schedule_multiplier(name, int(s), warm, total, G["decay_steps"].tolist(), G["decay_scales"].tolist())
               for s in G["steps"]]
        np.testing.assert_allclose(got, G[name], rtol=1e-13, atol=0, err_msg=name)


def test_optimizer_and_lambda_lr_wiring_match_reference():
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 3)
    cfg = OptimConfig(lr=0.24)
    opt = build_optimizer(model, cfg)
    g = opt.param_groups[0]
    np.testing.assert_array_equal([g["momentum"], g["weight_decay"], float(g["nesterov"]), g["dampening"]],
                                  G["sgd_group"])
    sched = build_scheduler(opt, iters_per_epoch=10, epochs=5, cfg=cfg)
    lrs = []
    for _ in range(50):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    np.testing.assert_allclose(lrs, G["lambda_lr"], rtol=1e-13)


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Linear(8, 8)
        self.classifier = torch.nn.Linear(8, 3)

    def forward(self, batch):
        logits = self.classifier(torch.relu(self.body(batch["x"])))
        return {"loss": torch.nn.functional.cross_entropy(logits, batch["y"])}


def test_train_step_order_clip_and_checkpoint_round_trip():
    torch.manual_seed(0)
    model = _Toy()
    cfg = OptimConfig(optimizer="sgd_fc", lr=0.5, grad_norm_clip=0.05)
    opt = build_optimizer(model, cfg)
    assert opt.param_groups[-1]["lr"] == 5.0 and len(opt.param_groups) == 3      # body w, body b, classifier
    sched = build_scheduler(opt, 4, 3, cfg)
    step = TrainStep(model, opt, sched, cfg.grad_norm_clip, amp=False, device_type="cpu")
    batch = {"x": torch.randn(16, 8), "y": torch.randint(0, 3, (16,))}
    before = [p.detach().clone() for p in model.parameters()]
    loss0 = step(batch)
    # first step: lr = base * min_scale (LambdaLR at step 0); the update is lr * (clipped grad + wd * p), so
    # the per-group displacement / lr has norm <= clip + wd * |p|
    disp = 0.0
    for grp in opt.param_groups:
        for p_new in grp["params"]:
            idx = [i for i, q in enumerate(model.parameters()) if q is p_new][0]
            disp += float(((p_new.detach() - before[idx]) / grp["lr"]).pow(2).sum())
    wd_norm = cfg.weight_decay * float(torch.sqrt(sum(b.pow(2).sum() for b in before)))
    assert abs(opt.param_groups[0]["lr"] / 0.5 - schedule_multiplier(cfg.scheduler, 1, 4, 12)) < 1e-12
    assert disp ** 0.5 <= cfg.grad_norm_clip + wd_norm + 1e-6
    for _ in range(5):
        step(batch)
    assert step.it == 6 and float(step(batch)) < float(loss0)
    buf = io.BytesIO()                                   # a real checkpoint round trip (no aliased buffers)
    torch.save(step.state_dict(), buf)
    buf.seek(0)
    state = torch.load(buf, weights_only=False)
    model2 = _Toy()
    opt2 = build_optimizer(model2, cfg)
    step2 = TrainStep(model2, opt2, build_scheduler(opt2, 4, 3, cfg), cfg.grad_norm_clip, amp=False, device_type="cpu")
    step2.load_state_dict(state)
    a, b = step(batch), step2(batch)
    assert torch.equal(a, b) and step2.it == step.it
    for p, q in zip(model.parameters(), model2.parameters()):
        assert torch.equal(p, q)
