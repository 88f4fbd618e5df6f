"""Drop-in for model/loss.py:22-115 (loss_per_scale) and the fused YOLOLayer training branch.

The constants the reference hard-codes (model/loss.py:35-41: gains 1/1/2, focal alpha .75/.5,
gamma 2) are compiled into the kernel; `opt` supplies stride, bbox_loss, ignore_thresh and
l1_loss_gain exactly as in the reference.
"""
from __future__ import annotations

import torch

from . import _ops, config


class _LossFn(torch.autograd.Function):
    """Forward computes the four losses AND d loss/d x in one kernel launch; backward only applies the
    upstream mix: the box / objectness / class channel groups carry d bbox_loss, d conf_loss,
    d cls_loss, so any upstream (g_loss, g_bbox, g_conf, g_cls) is a per-group scale, done in place
    by a kernel that exits immediately when all factors are 1 (no host sync either way)."""

    @staticmethod
    def forward(ctx, x, label, bboxes, num_classes, stride, bbox_loss, ignore_thresh, l1_gain, input_is_raw):
        want_grad = x.requires_grad
        out, flag, grad = _ops.loss_fwd_bwd(x.detach(), input_is_raw, label, bboxes, num_classes, stride,
                                            bbox_loss, ignore_thresh, l1_gain, want_grad)
        ctx.meta = (num_classes, input_is_raw)
        ctx.pq_grad = grad
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(flag)
        return out[0:1], out[1:2], out[2:3], out[3:4], flag

    @staticmethod
    def backward(ctx, g_loss, g_bbox, g_conf, g_cls, _g_flag):
        grad = ctx.pq_grad
        if grad is None:
            if any(g is not None for g in (g_loss, g_bbox, g_conf, g_cls)) and ctx.needs_input_grad[0]:
                raise RuntimeError("pqdet loss: backward called twice (the fused gradient is consumed "
                                   "in place by the first call)")
            return (None,) * 9
        ctx.pq_grad = None
        num_classes, input_is_raw = ctx.meta
        _ops.loss_scale_grad(grad, input_is_raw, num_classes, g_loss, g_bbox, g_conf, g_cls)
        return (grad,) + (None,) * 8


def _raise_if_nan(out4_views, flag):
    # model/loss.py:110-114
    if int(flag.item()) != 0:
        loss, lb, lc, lp = out4_views
        print('xy: {}, conf: {}, cls: {}'.format(lb.item(), lc.item(), lp.item()))
        raise RuntimeError('NaN in loss')


def _yolo_loss(x, label, bboxes, opt, input_is_raw: bool):
    stride = opt['stride']
    bbox_loss = opt['bbox_loss']
    if bbox_loss not in ('l1', 'giou', 'diou', 'ciou', 'iou'):
        raise NotImplementedError
    num_classes = label.shape[-1] - 6
    loss, lb, lc, lp, flag = _LossFn.apply(x, label, bboxes, num_classes, stride, bbox_loss,
                                           opt['ignore_thresh'], opt.get('l1_loss_gain', 0.1), input_is_raw)
    if config.nan_check == "sync":
        _raise_if_nan((loss, lb, lc, lp), flag)
    elif config.nan_check == "lazy":
        loss.pq_nan_flag = flag
    return loss, lb, lc, lp


def loss_per_scale(pred, label, bboxes, opt):
    """model/loss.py:22-115: pred (B,H,W,A,5+C) decoded, label (B,H,W,A,6+C), bboxes (B,G,4)
    -> (loss, bbox_loss, conf_loss, prob_loss), each shape (1,), autograd-connected to pred."""
    return _yolo_loss(pred, label, bboxes, opt, input_is_raw=False)
