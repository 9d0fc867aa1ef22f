"""Training-step driver of the segmentors: optimizer / LR-schedule builders and the
AMP step with the reference's semantics.

reference: pcseg/optim/__init__.py:13-71 (optimizers; note the reference never passes NESTEROV to SGD),
:74-112 (schedule multipliers), :115-170 (LambdaLR wiring), train.py:360-373 (zero_grad -> autocast
forward -> scaled backward -> unscale -> clip_grad_norm_ -> scaler.step -> scaler.update -> scheduler.step).
``bench.py`` times exactly this step.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Sequence

import torch
from torch import nn

__all__ = ["OptimConfig", "schedule_multiplier", "build_optimizer", "build_scheduler", "TrainStep"]


@dataclass
class OptimConfig:
    """tools/cfgs/**/*.yaml OPTIM block (defaults: voxel/semantic_kitti/minkunet_mk34_cr10.yaml:25-35)."""
    optimizer: str = "sgd"
    lr: float = 0.02 * 12                  # LR_PER_SAMPLE x BATCH_SIZE_PER_GPU (train.py:251)
    weight_decay: float = 1e-4
    momentum: float = 0.9
    betas: Sequence[float] = (0.9, 0.999)
    eps: float = 1e-8
    grad_norm_clip: float = 10.0
    scheduler: str = "linear_warmup_with_cosdecay"
    warmup_epoch: int = 1
    decay_epochs: Sequence[int] = field(default_factory=list)
    decay_scales: Sequence[float] = field(default_factory=list)


def schedule_multiplier(name: str, step: int, warmup_steps: int, total_steps: int,
                        decay_steps: Sequence[int] = (), decay_scales: Sequence[float] = (),
                        min_scale: float = 1e-5) -> float:
    """LR multiplier at ``step`` for the four named schedules.  As in the reference the cosine phase is
    measured against ``total_steps`` (not total - warmup), so the multiplier does not reach min_scale."""
    if name in ("linear_warmup_with_cosdecay", "cos_warmup_with_cosdecay"):
        if step < warmup_steps:
            ramp = step / warmup_steps if name.startswith("linear") else (1 - math.cos(math.pi * step / warmup_steps)) / 2
            return (1 - min_scale) * ramp + min_scale
        ratio = (step - warmup_steps) / total_steps
        return (1 - min_scale) * 0.5 * (1 + math.cos(math.pi * ratio)) + min_scale
    if name in ("linear_warmup_with_stepdecay", "coswarmup_with_stepdecay"):
        if step < warmup_steps:
            return step / warmup_steps if name.startswith("linear") else (1 - math.cos(math.pi * step / warmup_steps)) / 2
        scale = 1.0
        for at, by in zip(decay_steps, decay_scales):
            if step >= at:
                scale *= by
        return scale
    raise NotImplementedError(f"scheduler {name!r}")


def build_optimizer(model: nn.Module, cfg: OptimConfig) -> torch.optim.Optimizer:
    if cfg.optimizer == "sgd":
        return torch.optim.SGD(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay, momentum=cfg.momentum)
    if cfg.optimizer == "sgd_fc":                      # 10x learning rate on the classifier head
        groups = [{"params": p} for n, p in model.named_parameters() if "classifier" not in n]
        groups.append({"params": model.classifier.parameters(), "lr": cfg.lr * 10})
        return torch.optim.SGD(groups, lr=cfg.lr, weight_decay=cfg.weight_decay, momentum=cfg.momentum)
    if cfg.optimizer == "adam":
        return torch.optim.Adam(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay)
    if cfg.optimizer == "adamw":
        return torch.optim.AdamW(model.parameters(), lr=cfg.lr, betas=tuple(cfg.betas), weight_decay=cfg.weight_decay,
                                 eps=cfg.eps)
    raise NotImplementedError(f"optimizer {cfg.optimizer!r}")


def build_scheduler(optimizer, iters_per_epoch: int, epochs: int, cfg: OptimConfig):
    warmup, total = cfg.warmup_epoch * iters_per_epoch, epochs * iters_per_epoch
    decay_steps = [e * iters_per_epoch for e in cfg.decay_epochs]
    assert len(cfg.decay_scales) == len(cfg.decay_epochs), "DECAY_SCALES does not match DECAY_EPOCHS"
    fn: Callable[[int], float] = lambda s: schedule_multiplier(cfg.scheduler, s, warmup, total, decay_steps,
                                                               cfg.decay_scales)
    fn(0)                                              # unknown names fail here, not at the first step
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=fn)


class TrainStep:
    """One optimisation step with the reference's ordering (train.py:360-373)."""

    def __init__(self, model: nn.Module, optimizer, scheduler=None, grad_norm_clip: float = 10.0,
                 amp: bool = True, device_type: str = "cuda"):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.clip, self.amp, self.device_type = grad_norm_clip, amp, device_type
        self.scaler = torch.amp.GradScaler(device_type, enabled=amp and device_type == "cuda")
        self.it = 0

    def __call__(self, batch: Dict) -> torch.Tensor:
        self.model.train()
        self.optimizer.zero_grad()
        with torch.autocast(self.device_type, dtype=torch.float16 if self.device_type == "cuda" else torch.bfloat16,
                            enabled=self.amp):
            loss = self.model(batch)["loss"].mean()
        self.scaler.scale(loss).backward()
        self.scaler.unscale_(self.optimizer)
        nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.scaler.step(self.optimizer)
        self.scaler.update()
        if self.scheduler is not None:
            self.scheduler.step()
        self.it += 1
        return loss.detach()

    def state_dict(self) -> Dict:
        """The reference's checkpoint entries (train.py:283-300)."""
        return {"it": self.it, "model_state": self.model.state_dict(), "optimizer_state": self.optimizer.state_dict(),
                "scaler_state": self.scaler.state_dict(),
                "scheduler_state": self.scheduler.state_dict() if self.scheduler is not None else None}

    def load_state_dict(self, state: Dict) -> None:
        self.it = state["it"]
        self.model.load_state_dict(state["model_state"])
        self.optimizer.load_state_dict(state["optimizer_state"])
        self.scaler.load_state_dict(state["scaler_state"])
        if self.scheduler is not None and state.get("scheduler_state") is not None:
            self.scheduler.load_state_dict(state["scheduler_state"])
