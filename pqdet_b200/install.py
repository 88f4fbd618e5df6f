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
