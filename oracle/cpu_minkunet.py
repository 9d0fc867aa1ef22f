"""CPU MinkUNet forward+backward built on the reference's CPU path.  TEST / BASELINE ONLY.

Used by (a) ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm and (b) the
end-to-end parity test of the CUDA model.  It restates the *orchestration* of
``TS/nn/functional/conv.py:122-205`` and
``pcseg/model/segmentor/voxel/minkunet/{minkunet.py:385-422, utils.py:11-105}`` on CPU
tensors and executes every sparse op through ``oracle/_ref`` - the reference's own
compiled CPU backend (``convolution_forward_cpu`` etc., built by oracle/build_ref.py) -
when it is present; otherwise through the numpy restatement in ``ref_ops`` ("port").

Parameters are consumed from a ``state_dict`` with the reference's key names, so the
CUDA model and this CPU path can be run on identical weights.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as TF

from . import build_ref
from . import ref_ops as R

_REF = None
_REF_TRIED = False


def ref_backend():
    """The compiled reference CPU backend (pybind module) or None."""
    global _REF, _REF_TRIED
    if not _REF_TRIED:
        _REF_TRIED = True
        try:
            _REF = build_ref.load()
        except Exception:
            _REF = None
    return _REF


def kind() -> str:
    return "reference" if ref_backend() is not None else "port"


# ----------------------------------------------------------------- sparse ops on CPU
def _hash(coords: torch.Tensor, offsets: Optional[np.ndarray] = None) -> torch.Tensor:
    be = ref_backend()
    if be is not None and offsets is None:
        return be.hash_cpu(coords.contiguous())
    # kernel-hash: the reference CPU twin mishandles batch > 0 (hash_cpu.cpp:29) -> restatement
    return torch.from_numpy(R.sphash(coords.numpy(), offsets))


def _query(queries: torch.Tensor, refs: torch.Tensor) -> torch.Tensor:
    be = ref_backend()
    if be is not None:
        idx = torch.arange(refs.numel(), dtype=torch.long)
        return be.hash_query_cpu(queries.contiguous().view(-1), refs.contiguous(), idx).view(
            queries.shape) - 1
    return torch.from_numpy(R.sphashquery(queries.numpy(), refs.numpy()))


def build_kmap(in_coords: torch.Tensor, out_coords: torch.Tensor, ks, in_stride):
    off = R.get_kernel_offsets(ks, in_stride)
    res = _query(_hash(out_coords, off), _hash(in_coords))
    hit = res != -1
    nbsizes = hit.sum(1)
    nz = torch.nonzero(hit)
    nbmaps = torch.stack([res[hit], nz[:, 1]], 1)
    return nbmaps.int().contiguous(), nbsizes.int().contiguous(), (in_coords.shape[0], out_coords.shape[0])


class _Conv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, nbmaps, nbsizes, sizes, transposed):
        be = ref_backend()
        x, w = x.contiguous(), w.contiguous()
        out = torch.zeros(sizes[0] if transposed else sizes[1], w.shape[-1])
        if be is not None:
            be.convolution_forward_cpu(x, out, w, nbmaps, nbsizes, transposed)
        else:
            a = 0
            for k in range(w.shape[0]):
                b = a + int(nbsizes[k])
                i = nbmaps[a:b, 1 if transposed else 0].long()
                o = nbmaps[a:b, 0 if transposed else 1].long()
                out.index_add_(0, o, x[i] @ w[k])
                a = b
        ctx.save_for_backward(x, w, nbmaps, nbsizes)
        ctx.transposed = transposed
        return out

    @staticmethod
    def backward(ctx, go):
        x, w, nbmaps, nbsizes = ctx.saved_tensors
        be = ref_backend()
        gx, gw = torch.zeros_like(x), torch.zeros_like(w)
        if be is not None:
            be.convolution_backward_cpu(x, gx, go.contiguous(), w, gw, nbmaps, nbsizes, ctx.transposed)
        else:
            a = 0
            for k in range(w.shape[0]):
                b = a + int(nbsizes[k])
                i = nbmaps[a:b, 1 if ctx.transposed else 0].long()
                o = nbmaps[a:b, 0 if ctx.transposed else 1].long()
                gx.index_add_(0, i, go[o] @ w[k].t())
                gw[k] = x[i].t() @ go[o]
                a = b
        return gx, gw, None, None, None, None


class _Voxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, idx, cnt):
        ctx.aux = (idx, cnt, f.shape[0])
        be = ref_backend()
        if be is not None and bool((idx >= 0).all()):
            return be.voxelize_forward_cpu(f.contiguous(), idx, cnt)
        return torch.from_numpy(R.spvoxelize_forward(f.numpy(), idx.numpy(), cnt.numpy()))

    @staticmethod
    def backward(ctx, g):
        idx, cnt, n = ctx.aux
        return torch.from_numpy(R.spvoxelize_backward(g.numpy(), idx.numpy(), cnt.numpy(), n)), None, None


class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, idx, w):
        ctx.aux = (idx, w, f.shape[0])
        be = ref_backend()
        if be is not None:
            return be.devoxelize_forward_cpu(f.contiguous(), idx, w)
        return torch.from_numpy(R.spdevoxelize_forward(f.numpy(), idx.numpy(), w.numpy()))

    @staticmethod
    def backward(ctx, g):
        idx, w, n = ctx.aux     # the reference CPU twin of this op is wrong -> restatement
        return torch.from_numpy(R.spdevoxelize_backward(g.numpy(), idx.numpy(), w.numpy(), n)), None, None


# ------------------------------------------------------------------------- the network
class CpuMinkUNet:
    """Functional MinkUNet (ResBlock variant) over a reference-named state_dict."""

    def __init__(self, state: Dict[str, torch.Tensor], num_layer=(2, 3, 4, 6, 2, 2, 2, 2),
                 pres: float = 0.05, vres: float = 0.05, ignore_label: int = 0,
                 label_smoothing: float = 0.1):
        self.p = {k: v.detach().float().cpu().clone().requires_grad_(v.is_floating_point() and
                                                                      "running" not in k)
                  for k, v in state.items()}
        self.num_layer, self.pres, self.vres = num_layer, pres, vres
        self.ignore_label, self.label_smoothing = ignore_label, label_smoothing

    # -- layer helpers; every tensor bundle is (feats, coords, stride)
    def _bn(self, x, name):
        return TF.batch_norm(x, None, None, self.p[name + ".weight"], self.p[name + ".bias"], True, 0.1, 1e-5)

    def _conv(self, x, name, ks, maps, stride=1, transposed=False):
        feats, coords, ts = x
        w = self.p[name + ".kernel"]
        if ks == 1:
            return (feats @ w, coords, ts)
        cm, km = maps
        if not transposed:
            out_ts = ts * stride
            if out_ts not in cm:
                cm[out_ts] = torch.from_numpy(R.spdownsample(coords.numpy(), stride, ks, ts))
            oc = cm[out_ts]
            key = (ts, ks, stride)
            if key not in km:
                km[key] = build_kmap(coords, oc, ks, ts)
        else:
            out_ts = ts // stride
            oc = cm[out_ts]
            key = (out_ts, ks, stride)
        return (_Conv.apply(feats, w, *km[key], transposed), oc, out_ts)

    def _block(self, x, name, ks, maps, stride=1, transposed=False):
        y = self._conv(x, name + ".net.0", ks, maps, stride, transposed)
        return (torch.relu(self._bn(y[0], name + ".net.1")), y[1], y[2])

    def _res(self, x, name, maps):
        y = self._conv(x, name + ".net.0", 3, maps)
        y = (torch.relu(self._bn(y[0], name + ".net.1")), y[1], y[2])
        y = self._conv(y, name + ".net.3", 3, maps)
        main = self._bn(y[0], name + ".net.4")
        if name + ".downsample.0.kernel" in self.p:
            sc = self._bn(x[0] @ self.p[name + ".downsample.0.kernel"], name + ".downsample.1")
        else:
            sc = x[0]
        return (torch.relu(main + sc), y[1], y[2])

    def _v2p(self, x, pts, cache):
        feats, coords, ts = x
        if ts not in cache:
            idx, w = R.trilinear_map(pts.numpy(), coords.numpy(), ts)
            cache[ts] = (torch.from_numpy(idx).int(), torch.from_numpy(w))
        return _Devoxelize.apply(feats, *cache[ts])

    def forward(self, coords: torch.Tensor, feats: torch.Tensor, labels: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        maps = ({}, {})
        vc, _, idxq, cnt, nfc = R.initial_voxelize(coords.float().numpy(), feats.numpy(), self.pres,
                                                   self.vres)
        idxq_t, cnt_t = torch.from_numpy(idxq).int(), torch.from_numpy(cnt)
        x = (_Voxelize.apply(feats, idxq_t, cnt_t), torch.from_numpy(vc), 1)
        maps[0][1] = x[1]
        pts, cache = torch.from_numpy(nfc), {}
        for i in (0, 3):
            y = self._conv(x, f"stem.{i}", 3, maps)
            x = (torch.relu(self._bn(y[0], f"stem.{i + 1}")), y[1], y[2])
        x0 = x
        z0 = self._v2p(x0, pts, cache)
        enc = [x0]
        for s in range(4):
            x = self._block(x, f"stage{s + 1}.0", 2, maps, stride=2)
            for j in range(self.num_layer[s]):
                x = self._res(x, f"stage{s + 1}.{j + 1}", maps)
            enc.append(x)
        z1 = self._v2p(enc[4], pts, cache)
        zs = [z1]
        for u in range(4):
            y = self._block(x, f"up{u + 1}.0", 2, maps, stride=2, transposed=True)
            skip = enc[3 - u]
            x = (torch.cat([y[0], skip[0]], 1), y[1], y[2])
            for j in range(self.num_layer[4 + u]):
                x = self._res(x, f"up{u + 1}.1.{j}", maps)
            if u in (1, 3):
                zs.append(self._v2p(x, pts, cache))
        logits = torch.cat(zs, 1) @ self.p["classifier.0.weight"].t() + self.p["classifier.0.bias"]
        loss = None
        if labels is not None:
            loss = self.loss(logits, labels.long())
        self.aux = {"z0": z0}
        return logits, loss

    def loss(self, logits, target):
        ce = TF.cross_entropy(logits, target, ignore_index=self.ignore_label,
                              label_smoothing=self.label_smoothing)
        probs = logits.softmax(1)
        keep = target != self.ignore_label
        p, t = probs[keep], target[keep]
        terms = []
        for c in range(probs.shape[1]):                      # lovasz_losses.py:174-203, 'present'
            fg = (t == c).float()
            if fg.sum() == 0:
                continue
            err = (fg - p[:, c]).abs()
            err_s, perm = torch.sort(err, 0, descending=True)
            fg_s = fg[perm]
            gts = fg_s.sum()
            inter = gts - fg_s.cumsum(0)
            union = gts + (1 - fg_s).cumsum(0)
            jac = 1.0 - inter / union
            jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
            terms.append(torch.dot(err_s, jac))
        lov = torch.stack(terms).mean() if terms else probs.sum() * 0
        return ce + lov

    def grads(self) -> Dict[str, torch.Tensor]:
        return {k: v.grad for k, v in self.p.items() if v.requires_grad and v.grad is not None}
