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


def install(strict: bool = False) -> dict:
    """Patch every reference module that is importable; returns {attribute path: True/False}."""
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
    return done
