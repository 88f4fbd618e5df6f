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
from .train_dataset import SparseTarget


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


class _MultiLossFn(torch.autograd.Function):
    """All levels' decode + loss forward/backward in one kernel launch (csrc/loss.cu loss_levels_kernel).
    Inputs: L raw heads, then L labels, then L GT lists.  Output: one (4+5L,) tensor (see the header)."""

    @staticmethod
    def forward(ctx, L, num_classes, strides, bbox_loss, ignore_thresh, l1_gain, *tensors):
        raws, labels, gts = tensors[:L], tensors[L:2 * L], tensors[2 * L:3 * L]
        want_grad = any(r.requires_grad for r in raws)
        out, flag, grads = _ops.loss_levels([r.detach() for r in raws], labels, gts, num_classes, strides,
                                            bbox_loss, ignore_thresh, l1_gain, want_grad)
        ctx.pq = (L, num_classes, grads)
        ctx.mark_non_differentiable(flag)
        return out, flag

    @staticmethod
    def backward(ctx, g_out, _g_flag):
        L, num_classes, grads = ctx.pq
        if grads is None:
            raise RuntimeError("pqdet loss: backward called twice (the fused gradient is consumed in place)")
        ctx.pq = (L, num_classes, None)
        _ops.loss_levels_scale_grad(grads, num_classes, g_out.contiguous())
        return (None,) * 6 + tuple(grads) + (None,) * (2 * L)


class _MultiLossSparseFn(torch.autograd.Function):
    """_MultiLossFn on a SparseTarget: inputs are L raw heads, then gt6, L owner maps, L GT lists."""

    @staticmethod
    def forward(ctx, L, num_classes, strides, bbox_loss, ignore_thresh, l1_gain, *tensors):
        raws, gt6 = tensors[:L], tensors[L]
        owners, gts = tensors[L + 1:2 * L + 1], tensors[2 * L + 1:3 * L + 1]
        want_grad = any(r.requires_grad for r in raws)
        out, flag, grads = _ops.loss_levels_sparse([r.detach() for r in raws], owners, gt6, gts, num_classes,
                                                   strides, bbox_loss, ignore_thresh, l1_gain, want_grad)
        ctx.pq = (L, num_classes, grads)
        ctx.mark_non_differentiable(flag)
        return out, flag

    @staticmethod
    def backward(ctx, g_out, _g_flag):
        L, num_classes, grads = ctx.pq
        if grads is None:
            raise RuntimeError("pqdet loss: backward called twice (the fused gradient is consumed in place)")
        ctx.pq = (L, num_classes, None)
        _ops.loss_levels_scale_grad(grads, num_classes, g_out.contiguous())
        return (None,) * 6 + tuple(grads) + (None,) * (2 * L + 1)


_SCALE_SLOT = {8: 0, 16: 1, 32: 2}



def combine_eval(layers, outs: Sequence[torch.Tensor]) -> torch.Tensor:
    """The eval branch of DetectionModel.forward (model/interpreter.py:72-76) on what the [yolo] layers RECEIVED: raw
    heads, or - after install.fuse_head_convs - the inputs of the head convolutions, tagged with their nn.Conv2d.
    Every level is decoded straight into its row range of the (B, N, 5+C) prediction: no per-level tensor, no cat."""
    C = layers[0].opt['classes']
    ch = 5 + C
    strides = [l.opt['stride'] for l in layers]
    convs = [getattr(o, '_pq_pending_conv', None) for o in outs]
    if torch.is_grad_enabled() and any(o.requires_grad for o in outs):
        # someone differentiates through an eval-mode prediction (saliency maps, adversarial examples,
        # distillation): the one-launch concat has no grad_fn, so take the per-level autograd-aware Decode
        # (model/interpreter.py:72-76 as written).  _FusedHeadConv never tags its input in this situation.
        dec = [l.decode(o) for l, o in zip(layers, outs)]
        return torch.cat([d.view((d.shape[0], -1, d.shape[-1])) for d in dec], dim=1)
    if all(c is None for c in convs):
        return _ops.decode_levels(list(outs), C, strides)
    B = outs[0].shape[0]
    A = [(o.shape[1] if c is None else c.out_channels) // ch for o, c in zip(outs, convs)]
    rows = [o.shape[2] * o.shape[3] * a for o, a in zip(outs, A)]
    N = sum(rows)
    out = torch.empty((B, N, ch), dtype=torch.float32, device=outs[0].device)
    if B and all(c is not None for c in convs):
        tiles = sum(B * ((o.shape[2] * o.shape[3] + 127) // 128) for o in outs)
        if (all((o.shape[2] * o.shape[3]) % 128 == 0 for o in outs) and tiles <= 32 * _sm_count(outs[0].device)
                and _ops.head_conv_decode_levels(outs, [c.weight for c in convs], [c.bias for c in convs], C, strides, out)):
            return out
    off = 0
    for o, c, s, r in zip(outs, convs, strides, rows):
        if c is None:
            _ops.decode_fwd(o, C, s, out=out, rows_total=N, row_offset=off)
        else:
            _ops.head_conv_decode(o, c.weight, c.bias, C, s, out=out, rows_total=N, row_offset=off)
        off += r
    return out


def combine_train(layers, outs: Sequence[torch.Tensor], target):
    """The training branch of DetectionModel.forward (model/interpreter.py:77-85) on what the [yolo] layers RECEIVED
    (raw heads): decode + loss + gradient of every level in ONE launch instead of one YOLOLayer call per level plus
    the Python sums.  Same dict, same values (the per-level kernels and the multi-level kernel share their code)."""
    raws = []
    for o in outs:
        conv = getattr(o, '_pq_pending_conv', None)        # eval-mode validation loss behind fuse_head_convs
        raws.append(o if conv is None else torch.nn.functional.conv2d(o, conv.weight, conv.bias))
    return multi_level_loss(layers, raws, target)


_SM_COUNT = {}


def _sm_count(device: torch.device) -> int:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def multi_level_loss(layers, heads: Sequence[torch.Tensor], target):
    """All levels' decode + loss (+ gradient) in one launch.  layers: the YOLOLayers in cfg order; heads: their raw
    inputs; target: the reference's 6-tuple (model/interpreter.py:16-20) or a train_dataset.SparseTarget."""
    L = len(layers)
    opt0 = layers[0].opt
    if isinstance(target, SparseTarget):
        same_opts = all(l.opt['bbox_loss'] == opt0['bbox_loss'] and l.opt['ignore_thresh'] == opt0['ignore_thresh']
                        and l.opt.get('l1_loss_gain', 0.1) == opt0.get('l1_loss_gain', 0.1) for l in layers)
        if not same_opts:
            raise ValueError("sparse targets need the same loss options on every [yolo] level")
        if opt0['bbox_loss'] not in ('l1', 'giou', 'diou', 'iou', 'ciou'):
            raise NotImplementedError
        slots = [_SCALE_SLOT[l.opt['stride']] for l in layers]
        out, flag = _MultiLossSparseFn.apply(L, target.num_classes, [l.opt['stride'] for l in layers],
                                             opt0['bbox_loss'], opt0['ignore_thresh'],
                                             opt0.get('l1_loss_gain', 0.1), *heads, target.gt,
                                             *[target.owner[i] for i in slots], *[target.bboxes[i] for i in slots])
        return DetectionHead._result(out, flag, L)
    same = all(l.opt['bbox_loss'] == opt0['bbox_loss'] and l.opt['ignore_thresh'] == opt0['ignore_thresh']
               and l.opt.get('l1_loss_gain', 0.1) == opt0.get('l1_loss_gain', 0.1) for l in layers)
    if not same or opt0['bbox_loss'] not in ('l1', 'giou', 'diou', 'iou', 'ciou'):
        return DetectionHead._forward_per_level(layers, heads, target)
    pairs = [_TARGET_MAP[l.opt['stride']](target) for l in layers]
    C = pairs[0][0].shape[-1] - 6
    out, flag = _MultiLossFn.apply(L, C, [l.opt['stride'] for l in layers], opt0['bbox_loss'],
                                   opt0['ignore_thresh'], opt0.get('l1_loss_gain', 0.1),
                                   *heads, *[p[0] for p in pairs], *[p[1] for p in pairs])
    return DetectionHead._result(out, flag, L)


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
            # eval: every level decoded straight into its row range of (B, N, 5+C): no cat pass, one launch
            C = self.layers[0].opt['classes']
            if all(l.opt['classes'] == C for l in self.layers) and len(heads) <= 4:
                return _ops.decode_levels(heads, C, [l.opt['stride'] for l in self.layers])
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
        return multi_level_loss(self.layers, heads, target)

    def forward_from_features(self, features: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                              biases: Sequence[torch.Tensor]) -> torch.Tensor:
        """Eval branch starting one layer earlier: features[l] = the input of level l's 1x1 head
        convolution (`filters = A*(5+C), size = 1, activation = linear`, model/cfg/*.cfg), weights/biases = that
        convolution's parameters.  Convolution (TF32 tensor cores, like PyTorch's own default) + Decode + the concat
        of model/interpreter.py:75-76 without the raw head ever being written: -> (B, N, 5+C)."""
        C = self.layers[0].opt['classes']
        ch = 5 + C
        B = features[0].shape[0]
        rows = [f.shape[2] * f.shape[3] * (w.shape[0] // ch) for f, w in zip(features, weights)]
        N = sum(rows)
        out = torch.empty((B, N, ch), dtype=torch.float32, device=features[0].device)
        # one launch for all levels when every level qualifies for the persistent kernel and the batch is small enough
        # for the launch gaps to matter (the C side applies the same rule; checking here saves marshalling twice)
        tiles = sum(B * ((f.shape[2] * f.shape[3] + 127) // 128) for f in features)
        if (B and all((f.shape[2] * f.shape[3]) % 128 == 0 for f in features)
                and tiles <= 32 * _sm_count(features[0].device)
                and _ops.head_conv_decode_levels(features, weights, biases, C, [l.opt['stride'] for l in self.layers], out)):
            return out
        off = 0
        for l, f, w, bi, r in zip(self.layers, features, weights, biases, rows):
            _ops.head_conv_decode(f, w, bi, C, l.opt['stride'], out=out, rows_total=N, row_offset=off)
            off += r
        return out

    def detect_from_features(self, features: Sequence[torch.Tensor], weights: Sequence[torch.Tensor], biases,
                             input_size, batch_original_size, dataset: str = "voc", score_threshold: float = 0.1,
                             iou_threshold: float = 0.45, **kw):
        """The whole eval post-process from the inputs of the head convolutions: convolution on the
        tensor cores whose epilogue thresholds, then decode + recover + class-aware NMS on the surviving rows
        (fused.features_nms).  -> fused.Detections, identical to fused.decode_nms on the raw heads."""
        from . import fused
        return fused.features_nms(features, weights, biases, [l.opt['stride'] for l in self.layers],
                                  self.layers[0].opt['classes'], input_size, batch_original_size, dataset,
                                  score_threshold, iou_threshold, **kw)

    def loss_and_grad(self, heads: Sequence[torch.Tensor], target):
        """The training branch without autograd glue: ONE kernel launch gives the loss dict of forward(heads,
        target) and d loss.mean() / d head for every level (the kernel always computes both).  A trainer continues
        into the backbone with torch.autograd.backward(list(heads), grads) instead of loss.mean().backward()
        (trainer.py:233) -- same gradients, minus ~8 tiny autograd kernels per step.
        target: the reference's 6-tuple or a train_dataset.SparseTarget.  -> (dict, [grad per level])"""
        L = len(self.layers)
        opt0 = self.layers[0].opt
        if not all(l.opt['bbox_loss'] == opt0['bbox_loss'] and l.opt['ignore_thresh'] == opt0['ignore_thresh']
                   and l.opt.get('l1_loss_gain', 0.1) == opt0.get('l1_loss_gain', 0.1) for l in self.layers):
            raise ValueError("loss_and_grad needs the same loss options on every [yolo] level")
        if opt0['bbox_loss'] not in ('l1', 'giou', 'diou', 'iou', 'ciou'):
            raise NotImplementedError
        strides = [l.opt['stride'] for l in self.layers]
        raws = [h.detach() for h in heads]
        if isinstance(target, SparseTarget):
            slots = [_SCALE_SLOT[s] for s in strides]
            out, flag, grads = _ops.loss_levels_sparse(raws, [target.owner[i] for i in slots], target.gt,
                                                       [target.bboxes[i] for i in slots], target.num_classes, strides,
                                                       opt0['bbox_loss'], opt0['ignore_thresh'],
                                                       opt0.get('l1_loss_gain', 0.1), True)
        else:
            pairs = [_TARGET_MAP[s](target) for s in strides]
            C = pairs[0][0].shape[-1] - 6
            out, flag, grads = _ops.loss_levels(raws, [p[0] for p in pairs], [p[1] for p in pairs], C, strides,
                                                opt0['bbox_loss'], opt0['ignore_thresh'],
                                                opt0.get('l1_loss_gain', 0.1), True)
        return self._result(out, flag, L), grads

    @staticmethod
    def _result(out, flag, L):
        if config.nan_check == "sync" and int(flag.item()) != 0:
            for l in range(L):                                        # model/loss.py:110-114
                lo = out[4 + 4 * l:8 + 4 * l]
                if bool(torch.isnan(lo[0])):
                    print('xy: {}, conf: {}, cls: {}'.format(lo[1].item(), lo[2].item(), lo[3].item()))
            raise RuntimeError('NaN in loss')
        res = {
            'loss': out[0:1],
            'giou_loss': out[1:2],
            'conf_loss': out[2:3],
            'class_loss': out[3:4],
            'loss_per_branch': [out[4 + 4 * L + l:5 + 4 * L + l] for l in range(L)],
        }
        if config.nan_check == "lazy":
            res['loss'].pq_nan_flag = flag
        res['loss'].pq_out = out              # the whole result vector, for dist.reduce_losses (one collective)
        return res

    @staticmethod
    def _forward_per_level(layers, heads, target):
        """One launch per level (levels with different loss options)."""
        mode = config.nan_check
        if mode == "sync":
            config.nan_check = "lazy"            # one host sync per step instead of one per level
        try:
            outputs = [l(h, _TARGET_MAP[l.opt['stride']](target)) for l, h in zip(layers, heads)]
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
