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


def fuse_head_convs(model) -> int:
    """SURVEY 8f-2 as a hook: let every `[yolo]` level of a reference DetectionModel (built after install(), weights
    loaded) run its 1x1 head convolution + Decode as ONE tensor-core kernel in eval mode.

    For each pqdet_b200.parser.YOLOLayer in `model.module_list` whose predecessor is the plain head convolution of the
    cfgs (`filters=A*(5+C), size=1, stride=1, activation=linear`, no batch norm: an nn.Sequential holding only `conv`,
    model/parser.py:385-410) and whose output no route / shortcut reads, the conv block's forward is overridden on
    the instance: in eval mode on CUDA it hands its INPUT to the YOLOLayer together with its own parameters
    (pqdet_head_conv_decode); in training mode, on CPU, for ONNX export or when a target is given it still convolves.
    No module is added or renamed, so state_dict keys, pruning and checkpoint loading are unchanged.  Returns the
    number of levels fused."""
    import types

    import torch

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
        conv = getattr(prev, 'conv', None)
        if getattr(prev, '_type', None) != 'convolutional' or not isinstance(conv, torch.nn.Conv2d) or len(prev) != 1:
            continue
        if (conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.padding != (0, 0) or conv.dilation != (1, 1)
                or conv.groups != 1 or conv.out_channels % (5 + layer.opt['classes']) or (i - 1) in used):
            continue

        def forward(self, x, _yolo=layer, _plain=type(prev).forward):
            if self.training or not x.is_cuda or _yolo.decode.onnx:
                return _plain(self, x)
            object.__setattr__(_yolo, '_pq_pending_conv', self.conv)      # not a submodule: state_dict unchanged
            return x

        prev.forward = types.MethodType(forward, prev)
        fused += 1
    return fused


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
