"""Drop-in for the `[yolo]` layer of the reference: model/parser.py:194-249 (Decode, YOLOLayer).

Same constructor arguments, same forward signatures, same output layout; the arithmetic runs in
the sm_100a kernels of csrc/decode.cu and csrc/loss.cu.  The training branch does NOT materialise
the decoded tensor: decode, loss and d loss/d conv are one fused pass over the raw head.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _ops
from .loss import _yolo_loss


def build_center_grid(height: int, width: int) -> torch.Tensor:
    """model/parser.py:185-192.  Kept for API completeness only: the kernels derive the cell centre
    (x + 0.5, y + 0.5) from the thread index, so no grid tensor is ever cached or respawned."""
    ys = torch.arange(0, height, dtype=torch.float32) + 0.5
    xs = torch.arange(0, width, dtype=torch.float32) + 0.5
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx.unsqueeze(-1), gy.unsqueeze(-1)], dim=-1)


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv, num_classes, stride):
        conv = conv.contiguous()
        ctx.save_for_backward(conv)
        ctx.meta = (num_classes, stride)
        return _ops.decode_fwd(conv, num_classes, stride)

    @staticmethod
    def backward(ctx, grad_out):
        (conv,) = ctx.saved_tensors
        num_classes, stride = ctx.meta
        return _ops.decode_bwd(conv, grad_out.contiguous(), num_classes, stride), None, None


class Decode(nn.Module):
    """model/parser.py:194-235: conv (B, A*(5+C), H, W) -> (B, H, W, A, 5+C)."""

    def __init__(self, num_classes: int, stride: int, onnx: bool = False):
        super().__init__()
        self.num_classes = num_classes
        self.stride = stride
        self.onnx = onnx          # accepted for signature compatibility; there is no grid cache to bypass

    def forward(self, conv: torch.Tensor) -> torch.Tensor:
        if conv.requires_grad and torch.is_grad_enabled():
            return _DecodeFn.apply(conv, self.num_classes, self.stride)
        return _ops.decode_fwd(conv, self.num_classes, self.stride)


class YOLOLayer(nn.Module):
    """model/parser.py:237-249.  opt = the cfg's [yolo] options + computed 'stride'
    (model/parser.py:453-459); target = (label, bboxes) or None."""

    def __init__(self, opt: dict, onnx: bool = False):
        super().__init__()
        self.decode = Decode(opt['classes'], opt['stride'], onnx)
        self.opt = opt

    def forward(self, x, target=None):
        if self.__dict__.get('_pq_passthrough'):
            # install.fuse_eval_concat / fuse_train_levels: the model combines all levels in one launch
            return x
        conv = getattr(x, '_pq_pending_conv', None)
        if conv is not None:
            # install.fuse_head_convs: x is the INPUT of this level's 1x1 head convolution (the conv block passed its
            # input through, tagged); eval runs convolution + Decode as one tensor-core kernel, anything else applies
            # the convolution first
            if target is None and not self.decode.onnx:
                return _ops.head_conv_decode(x, conv.weight, conv.bias, self.opt['classes'], self.opt['stride'])
            x = torch.nn.functional.conv2d(x, conv.weight, conv.bias)
        if target is None:
            return self.decode(x)
        label, bboxes = target
        return _yolo_loss(x, label, bboxes, self.opt, input_is_raw=True)
