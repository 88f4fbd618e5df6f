"""Monkey-patch the hot path into an importable copy of the reference (never edits its files).

    import tools                      # the reference's modules, imported by the user as usual
    import pqdet_b200.install
    pqdet_b200.install.install()

After this, DetectionModel built from any darknet cfg decodes / computes its loss with the sm_100a
kernels, Evaluator / predict.py recover boxes and run NMS with them, and TrainDataset.create_label
is available on the GPU.  The hook points are the ones SURVEY.md section 8b lists.
"""
from __future__ import annotations

import importlib
import sys

import torch
from torch import nn

from . import base_sample, loss, parser, tools as pq_tools


def install(strict: bool = False, patch_augment: bool = False) -> dict:
    """Patch every reference module that is importable; returns {attribute path: True/False}.
    patch_augment: also replace dataset.augment.Resize by the GPU letterbox.  Off by default: DataLoader worker
    processes (where the training pipeline calls Resize) must not touch CUDA; turn it on for single-process eval /
    predict scripts, or call pqdet_b200.augment.letterbox_normalize on whole batches instead."""
    done = {}

    def patch(modname, attr, value):
        key = "%s.%s" % (modname, attr)
        try:
            mod = sys.modules.get(modname) or importlib.import_module(modname)
            setattr(mod, attr, value)
            done[key] = True
        except Exception:
            if strict:
                raise
            done[key] = False

    for name in ("torch_nms", "iou_calc3", "giou", "diou", "ciou"):
        patch("tools", name, getattr(pq_tools, name))
    patch("model.loss", "loss_per_scale", loss.loss_per_scale)
    patch("model.parser", "Decode", parser.Decode)
    patch("model.parser", "YOLOLayer", parser.YOLOLayer)
    patch("model.parser", "loss_per_scale", loss.loss_per_scale)
    patch("dataset.base_sample", "recover_bboxes_prediction", base_sample.recover_bboxes_prediction)
    for ds in ("voc", "coco", "visdrone"):
        fn = base_sample.RECOVER_BBOXES_REGISTER[ds]
        patch("dataset.%s_sample" % ds, "recover_bboxes_prediction_%s" % ds, fn)
        try:
            importlib.import_module("dataset").RECOVER_BBOXES_REGISTER[ds] = fn
            done["dataset.RECOVER_BBOXES_REGISTER[%s]" % ds] = True
        except Exception:
            if strict:
                raise
            done["dataset.RECOVER_BBOXES_REGISTER[%s]" % ds] = False
    # the steps either side of the path: evaluator statistics and the eval letterbox
    try:
        _patch_evaluator()
        done["eval.evaluator.Evaluator.{init_statics,add_detections,add_labels,AP}"] = True
    except Exception:
        if strict:
            raise
        done["eval.evaluator.Evaluator.{init_statics,add_detections,add_labels,AP}"] = False
    if patch_augment:
        from . import augment as pq_augment
        patch("dataset.augment", "Resize", pq_augment.Resize)
    return done


class _FusedHeadConv(nn.Sequential):
    """Class a head-convolution block is switched to by fuse_head_convs (same attributes, same state_dict keys).  In
    CUDA eval mode it does not convolve: it passes its input on, tagged with its own nn.Conv2d, and the YOLOLayer that
    receives the tagged tensor runs convolution + Decode as one kernel.  No reference to another module is kept, so
    nn.DataParallel replicas (tools.py:215-216) work on their own parameters."""

    def forward(self, x):
        if self.training or not x.is_cuda or torch.onnx.is_in_onnx_export():
            return super().forward(x)
        y = x.view_as(x)                      # a new tensor object on the same storage: the tag stays private
        y._pq_pending_conv = self.conv
        return y


def fuse_head_convs(model) -> int:
    """SURVEY 8f-2 as a hook: let every `[yolo]` level of a reference DetectionModel (built after install()) run its
    1x1 head convolution + Decode as ONE tensor-core kernel in eval mode.

    For each pqdet_b200.parser.YOLOLayer in `model.module_list` whose predecessor is the plain head convolution of the
    cfgs (`filters=A*(5+C), size=1, stride=1, activation=linear`, no batch norm: an nn.Sequential holding only `conv`,
    model/parser.py:385-410) and whose output no route / shortcut reads, the conv block's class is switched to
    _FusedHeadConv: in eval mode on CUDA it hands its INPUT to the YOLOLayer together with its own parameters
    (pqdet_head_conv_decode); in training mode, on CPU, for ONNX export or when a target is given the convolution
    still runs.  No module is added or renamed, so state_dict keys, pruning and checkpoint loading are unchanged.
    Returns the number of levels fused."""
    layers = list(model.module_list)
    used = set()
    for j, layer in enumerate(layers):              # who reads which cached output (model/interpreter.py:46-50)
        t = getattr(layer, '_type', None)
        refs = []
        if t in ('shortcut', 'scale_channels'):
            refs = [layer._from]
        elif t == 'route':
            refs = list(layer._layers)
        for r in refs:
            used.add(r if r >= 0 else j + r)
    fused = 0
    for i, layer in enumerate(layers):
        if i == 0 or not isinstance(layer, parser.YOLOLayer) or getattr(model, 'quant', False):
            continue
        prev = layers[i - 1]
        if isinstance(prev, _FusedHeadConv):
            fused += 1
            continue
        conv = getattr(prev, 'conv', None)
        if (getattr(prev, '_type', None) != 'convolutional' or type(prev) is not nn.Sequential
                or not isinstance(conv, nn.Conv2d) or len(prev) != 1):
            continue
        if (conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.padding != (0, 0) or conv.dilation != (1, 1)
                or conv.groups != 1 or conv.out_channels % (5 + layer.opt['classes']) or (i - 1) in used):
            continue
        prev.__class__ = _FusedHeadConv
        fused += 1
    return fused


def fuse_eval_concat(model) -> bool:
    """SURVEY section 8 row a4 as a hook: the eval branch of DetectionModel.forward (model/interpreter.py:72-76: every
    [yolo] level decoded, viewed as (B, -1, 5+C) and concatenated) becomes one launch that decodes all levels
    straight into the (B, N, 5+C) prediction - no per-level tensor, no torch.cat (which alone moves as many bytes as
    the decode).  Combined with fuse_head_convs the head convolutions run inside that launch too.

    The model's class is switched to a subclass whose forward, in CUDA eval mode without a target, runs the parent
    class's layer loop (AnyModel.forward) with the YOLOLayers passing their input through and then combines the
    levels (interpreter.combine_eval).  Training, targets and CPU tensors take the reference's own forward.  Returns
    False (model untouched) if a [yolo] layer is not ours or something reads a [yolo] output."""
    from . import interpreter as pq_interpreter
    cls = type(model)
    if getattr(cls, '_pq_eval_concat', False):
        return True
    layers = list(model.module_list)
    yolo_idx = [i for i, l in enumerate(layers) if getattr(l, '_type', None) == 'yolo']
    if not yolo_idx or not all(isinstance(layers[i], parser.YOLOLayer) for i in yolo_idx) or getattr(model, 'quant', False):
        return False
    for j, layer in enumerate(layers):
        t = getattr(layer, '_type', None)
        refs = [layer._from] if t in ('shortcut', 'scale_channels') else list(layer._layers) if t == 'route' else []
        if any((r if r >= 0 else j + r) in yolo_idx for r in refs):
            return False
    loop = None
    for k in cls.__mro__[1:]:
        if 'forward' in k.__dict__:
            loop = k.__dict__['forward']          # AnyModel.forward: the layer loop, returns the [yolo] outputs
            break
    if loop is None:
        return False

    def forward(self, x, target=None):
        if target is not None or self.training or not x.is_cuda or torch.onnx.is_in_onnx_export():
            return cls.forward(self, x, target)
        yolos = [l for l in self.module_list if isinstance(l, parser.YOLOLayer)]
        for l in yolos:
            object.__setattr__(l, '_pq_passthrough', True)
        try:
            outs = loop(self, x, None)
        finally:
            for l in yolos:
                l.__dict__.pop('_pq_passthrough', None)
        if not isinstance(outs, (list, tuple)):
            outs = [outs]
        return pq_interpreter.combine_eval(yolos, outs)

    model.__class__ = type(cls.__name__, (cls,), {'forward': forward, '_pq_eval_concat': True})
    return True


def _patch_evaluator():
    """Evaluator keeps its loop (eval/evaluator.py:44-62); its statistics go through DetectionAccumulator, so AP()
    runs the matching on the GPU and returns the same tools.AP tuple."""
    ev_mod = importlib.import_module("eval.evaluator")
    ref_tools = sys.modules.get("tools") or importlib.import_module("tools")
    from .evaluator import DetectionAccumulator

    def init_statics(self):
        self._pq_acc = DetectionAccumulator(self._classes)
        self.detections_count = 0

    def add_detections(self, file_name, bboxes):
        self._pq_acc.add_detections(file_name, bboxes)
        self.detections_count = self._pq_acc.detections_count

    def add_labels(self, file_name, bboxes, diffs):
        self._pq_acc.add_labels(file_name, bboxes, diffs)

    def AP(self):
        m = self._pq_acc.AP()
        self.init_statics()
        return ref_tools.AP(m.mAPs, m.APs, m.AP, m.raw, m.class_names, m.iou_thresholds)

    for name, fn in (("init_statics", init_statics), ("add_detections", add_detections),
                     ("add_labels", add_labels), ("AP", AP)):
        setattr(ev_mod.Evaluator, name, fn)
