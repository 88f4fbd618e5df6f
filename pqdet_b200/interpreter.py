"""The part of model/interpreter.py that belongs to the hot path: how DetectionModel.forward
combines the per-level [yolo] outputs (model/interpreter.py:16-20, 72-85).

DetectionHead takes the raw head tensors the backbone/neck produce (stock PyTorch, out of scope)
in cfg order and returns exactly what DetectionModel.forward returns.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn

from . import _ops, config
from .loss import _raise_if_nan
from .parser import YOLOLayer


def item_getter(*items):
    def getter(x):
        if x is None:
            return None
        return tuple(x[i] for i in items)
    return getter


# model/interpreter.py:16-20: target = (label_s, label_m, label_l, sbboxes, mbboxes, lbboxes)
_TARGET_MAP = {
    8: item_getter(0, 3),
    16: item_getter(1, 4),
    32: item_getter(2, 5),
}


class DetectionHead(nn.Module):
    def __init__(self, opts: Sequence[dict], onnx: bool = False):
        """opts: one [yolo] option dict per level, in cfg order (FPN: strides 32, 16, 8)."""
        super().__init__()
        self.layers = nn.ModuleList([YOLOLayer(dict(o), onnx) for o in opts])

    def forward(self, heads: Sequence[torch.Tensor], target=None):
        if len(heads) != len(self.layers):
            raise ValueError("expected %d head tensors" % len(self.layers))
        if target is None:
            if any(h.requires_grad for h in heads) and torch.is_grad_enabled():
                outs = [l(h) for l, h in zip(self.layers, heads)]
                return torch.cat([o.view((o.shape[0], -1, o.shape[-1])) for o in outs], dim=1)
            # eval: decode every level straight into its row range of (B, N, 5+C) -- no cat pass
            C = self.layers[0].opt['classes']
            ch = 5 + C
            B = heads[0].shape[0]
            rows = [h.shape[2] * h.shape[3] * (h.shape[1] // ch) for h in heads]
            N = sum(rows)
            out = torch.empty((B, N, ch), dtype=torch.float32, device=heads[0].device)
            off = 0
            for l, h, r in zip(self.layers, heads, rows):
                _ops.decode_fwd(h, C, l.opt['stride'], out=out, rows_total=N, row_offset=off)
                off += r
            return out
        mode = config.nan_check
        if mode == "sync":
            config.nan_check = "lazy"            # one host sync per step instead of one per level
        try:
            outputs = [l(h, _TARGET_MAP[l.opt['stride']](target)) for l, h in zip(self.layers, heads)]
        finally:
            config.nan_check = mode
        if mode == "sync":
            flags = torch.cat([o[0].pq_nan_flag for o in outputs]).cpu()
            for o, f in zip(outputs, flags):
                if int(f):
                    _raise_if_nan(o, f)
        losses = list(map(sum, zip(*outputs)))
        loss_per_branch = [sum(loss[1:]) for loss in outputs]
        return {
            'loss': losses[0],
            'giou_loss': losses[1],
            'conf_loss': losses[2],
            'class_loss': losses[3],
            'loss_per_branch': loss_per_branch,
        }
