"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (pqdet_b200/).

Live-reference harness: imports the *unmodified* eleflea/PQDet sources from
/root/reference in-process (CPU tensors) so that

  * oracle/make_golden.py can generate the committed fixtures in tests/golden/, and
  * tests/test_oracle_vs_reference.py can pin oracle/pqdet_oracle.py against the
    reference itself on random inputs (skipped wherever /root/reference is absent,
    e.g. on the GPU box).

The reference has no tests/golden vectors of its own, so the
oracle's pin is "outputs of the reference itself run here".  Three harness-side shims
are needed; nothing under /root/reference is modified or copied:

  1. `tools` must be imported before `model.*` (tools.py:16-17 <-> model/loss.py:4 are circular);
  2. `yacs` is not installed: a stub CfgNode is injected (config.py:4);
  3. numpy >= 1.24 removed np.float/np.int/np.bool (dataset/train_dataset.py:126).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PQDET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "parser.py"))


_loaded = None


def load():
    """Import the reference and return a namespace with the hot-path callables."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "yacs" not in sys.modules:
        yacs, yc = types.ModuleType("yacs"), types.ModuleType("yacs.config")

        class CfgNode(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

        yc.CfgNode = CfgNode
        yacs.config = yc
        sys.modules["yacs"], sys.modules["yacs.config"] = yacs, yc
    for name, typ in (("float", float), ("int", int), ("bool", bool)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    warnings.filterwarnings("ignore", message="torch.meshgrid")
    import tools  # noqa: F401  (must precede model.*)
    from model.parser import Decode, YOLOLayer, build_center_grid
    from model.loss import loss_per_scale, smooth_l1_loss, focal
    from dataset import RECOVER_BBOXES_REGISTER
    from dataset.train_dataset import TrainDataset, collate_batch
    import config as ref_config

    ns = types.SimpleNamespace(
        tools=tools, Decode=Decode, YOLOLayer=YOLOLayer, build_center_grid=build_center_grid,
        loss_per_scale=loss_per_scale, smooth_l1_loss=smooth_l1_loss, focal=focal,
        RECOVER=RECOVER_BBOXES_REGISTER, TrainDataset=TrainDataset, collate_batch=collate_batch,
        config=ref_config,
    )
    _loaded = ns
    return ns


VOC_ANCHORS = [(10, 13), (16, 30), (33, 23), (30, 61), (62, 45), (59, 119),
               (116, 90), (156, 198), (373, 326)]          # config.py:58-59
VISDRONE_ANCHORS = [(9, 13), (25, 17), (16, 31), (47, 29), (32, 51), (83, 48),
                    (61, 91), (131, 99), (210, 189)]       # yamls/visdrone.yaml:16


def make_label_dataset(num_classes: int, anchors=VOC_ANCHORS, iou_threshold: float = 0.3):
    """A TrainDataset with only the fields create_label reads (dataset/train_dataset.py:110-146)."""
    ref = load()
    ds = ref.TrainDataset.__new__(ref.TrainDataset)
    ds._strides = np.array([8, 16, 32])
    ds._num_classes = num_classes
    ds._gt_per_grid = 3
    ds._anchors = np.array(anchors, dtype=np.float32)
    ds._anchors_iou_threshold = iou_threshold
    return ds


def make_evaluator(class_names):
    """The reference Evaluator with only the fields its statistics code reads (eval/evaluator.py:31-36, 64-175):
    no model, no dataset, no config."""
    load()
    from eval.evaluator import Evaluator
    ev = Evaluator.__new__(Evaluator)
    ev._classes = list(class_names)
    ev.init_statics()
    return ev
