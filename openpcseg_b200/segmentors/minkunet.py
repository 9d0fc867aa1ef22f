"""MinkUNet segmentor (the caller the benchmark drives), on the B200 sparse backend.

Architecture and state_dict key names follow the reference model so its checkpoints load
(pcseg/model/segmentor/voxel/minkunet/minkunet.py:186-458; config
tools/cfgs/voxel/semantic_kitti/minkunet_mk34_cr10.yaml): stem (2x k3) -> 4 stages of
[k2s2 down + n residual blocks] -> 4 up stages of [k2s2 transposed + skip concat + 2
residual blocks]; point features are sampled after the stem, stage 4, up 2 and up 4 and a
linear classifier runs on their concatenation.  Loss = cross-entropy (label smoothing) +
Lovasz-softmax, as in the reference (minkunet.py:344-362, 424-429).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import torch
from torch import nn

from .. import torchsparse as ts
from ..torchsparse import nn as spnn
from ..torchsparse.nn.functional import batch_norm_act, batch_norm_fusable, zero_sums
from ..torchsparse import PointTensor
from .losses import SegLoss
from ..torchsparse.operators import _CatFeats
from .point_voxel import initial_voxelize, voxel_to_point

__all__ = ["MinkUNet", "MinkUNetConfig", "minkunet34_config"]


@dataclass
class MinkUNetConfig:
    in_feature_dim: int = 4
    num_class: int = 20
    num_layer: Sequence[int] = (2, 3, 4, 6, 2, 2, 2, 2)
    planes: Sequence[int] = (32, 32, 64, 128, 256, 256, 128, 96, 96)
    cr: float = 1.0
    pres: float = 0.05
    vres: float = 0.05
    dropout_p: float = 0.0
    label_smoothing: float = 0.1
    ignore_label: int = 0
    sync_bn: bool = False
    channels: List[int] = field(init=False)

    def __post_init__(self):
        self.channels = [int(self.cr * c) for c in self.planes]


def minkunet34_config(**kw) -> MinkUNetConfig:
    """MinkUNet-34 cr1.0, ResBlock (minkunet_mk34_cr10.yaml:13-23)."""
    return MinkUNetConfig(**kw)


class _SparseBN(nn.BatchNorm1d):
    def forward(self, x):
        return x._like(super().forward(x.feats))


class _SparseSyncBN(nn.SyncBatchNorm):
    def forward(self, x):
        return x._like(super().forward(x.feats))


def _conv_bn(conv, bn, x, relu=True, residual=None):
    """conv -> fused (batch norm [+ residual] [+ ReLU]) on the conv's output rows."""
    sums = None
    if batch_norm_fusable(torch.float16 if torch.is_autocast_enabled() else x.feats.dtype, bn):
        # the conv's epilogue accumulates the batch statistics of its own output rows (one [N, C] pass less)
        sums = zero_sums(bn.num_features, x.feats.device)
    y = conv(x, bn_sums=sums)
    return y._like(batch_norm_act(y.feats, bn, relu=relu, residual=residual))


def _bn(c: int, sync: bool) -> nn.Module:
    return _SparseSyncBN(c) if sync else _SparseBN(c)


class ConvBlock(nn.Module):
    """conv -> BN -> ReLU; ``transposed`` gives the up-sampling variant."""

    def __init__(self, inc, outc, ks=3, stride=1, transposed=False, sync=False):
        super().__init__()
        self.net = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks, stride=stride,
                                             transposed=transposed), _bn(outc, sync), spnn.ReLU(True))

    def forward(self, x):                     # net = (conv, bn, relu): BN + ReLU run fused
        return _conv_bn(self.net[0], self.net[1], x)


class ResidualBlock(nn.Module):
    expansion = 1

    def __init__(self, inc, outc, ks=3, sync=False):
        super().__init__()
        self.net = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks), _bn(outc, sync), spnn.ReLU(True),
                                 spnn.Conv3d(outc, outc, kernel_size=ks), _bn(outc, sync))
        if inc == outc:
            self.downsample = nn.Identity()
        else:
            self.downsample = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=1), _bn(outc, sync))
        self.relu = spnn.ReLU(True)

    def forward(self, x):                     # relu(bn(conv(relu(bn(conv x)))) + shortcut(x))
        y = _conv_bn(self.net[0], self.net[1], x)
        if isinstance(self.downsample, nn.Identity):
            shortcut = x.feats
        else:
            shortcut = _conv_bn(self.downsample[0], self.downsample[1], x, relu=False).feats
        return _conv_bn(self.net[3], self.net[4], y, relu=True, residual=shortcut)


def _res_layers(inc: int, outc: int, n: int, sync: bool) -> List[nn.Module]:
    return [ResidualBlock(inc if i == 0 else outc, outc, sync=sync) for i in range(n)]


class MinkUNet(nn.Module):
    def __init__(self, cfg: MinkUNetConfig):
        super().__init__()
        self.cfg = cfg
        cs, nl, sync = cfg.channels, cfg.num_layer, cfg.sync_bn
        self.stem = nn.Sequential(spnn.Conv3d(cfg.in_feature_dim, cs[0], kernel_size=3), _bn(cs[0], sync),
                                  spnn.ReLU(True),
                                  spnn.Conv3d(cs[0], cs[0], kernel_size=3), _bn(cs[0], sync), spnn.ReLU(True))
        width = cs[0]
        for i in range(4):                               # encoder: stage1..stage4
            stage = nn.Sequential(ConvBlock(width, width, ks=2, stride=2, sync=sync),
                                  *_res_layers(width, cs[i + 1], nl[i], sync))
            setattr(self, f"stage{i + 1}", stage)
            width = cs[i + 1]
        skips = [cs[3], cs[2], cs[1], cs[0]]
        for i in range(4):                               # decoder: up1..up4
            up = nn.ModuleList([
                ConvBlock(width, cs[5 + i], ks=2, stride=2, transposed=True, sync=sync),
                nn.Sequential(*_res_layers(cs[5 + i] + skips[i], cs[5 + i], nl[4 + i], sync))])
            setattr(self, f"up{i + 1}", up)
            width = cs[5 + i]
        self.classifier = nn.Sequential(nn.Linear(cs[4] + cs[6] + cs[8], cfg.num_class))
        self.dropout = nn.Dropout(cfg.dropout_p, True)
        self.criterion = SegLoss(ignore_index=cfg.ignore_label, label_smoothing=cfg.label_smoothing)

    def _up(self, block: nn.ModuleList, x, skip):
        return block[1](ts.cat([block[0](x), skip]))

    def forward_logits(self, lidar: ts.SparseTensor) -> torch.Tensor:
        feats = lidar.F[:, : self.cfg.in_feature_dim]
        z = PointTensor(feats, lidar.C.float())
        x0 = initial_voxelize(z, self.cfg.pres, self.cfg.vres)
        x0 = _conv_bn(self.stem[0], self.stem[1], x0)
        x0 = _conv_bn(self.stem[3], self.stem[4], x0)
        z0 = voxel_to_point(x0, z)
        x1 = self.stage1(x0)
        x2 = self.stage2(x1)
        x3 = self.stage3(x2)
        x4 = self.stage4(x3)
        z1 = voxel_to_point(x4, z0)
        x4.F = self.dropout(x4.F)
        y1 = self._up(self.up1, x4, x3)
        y2 = self._up(self.up2, y1, x2)
        z2 = voxel_to_point(y2, z1)
        y2.F = self.dropout(y2.F)
        y3 = self._up(self.up3, y2, x1)
        y4 = self._up(self.up4, y3, x0)
        z3 = voxel_to_point(y4, z2)
        return self.classifier(_CatFeats.apply(z1.F, z2.F, z3.F))

    def forward(self, batch_dict):
        logits = self.forward_logits(batch_dict["lidar"])
        if "targets" not in batch_dict:
            return {"logits": logits}
        target = batch_dict["targets"]
        target = target.F if hasattr(target, "F") else target
        return {"loss": self.criterion(logits, target.long().view(-1)), "logits": logits}
