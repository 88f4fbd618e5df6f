"""Drop-in for dataset/base_sample.py:98-139 and the per-dataset wrappers
(voc_sample.py:97-104, coco_sample.py:102-109, visdrone_sample.py:90-97)."""
from __future__ import annotations

import torch

from . import _ops


def recover_bboxes_prediction(batch_pred_bbox: torch.Tensor, input_size, batch_original_size,
                              affine_func) -> torch.Tensor:
    """(B, N, 5+C) decoded -> (B, N, 4+C) in original-image coordinates, clipped, conf folded into
    the class scores.  `affine_func`: one of the three affines the reference ships - this module's tags, the
    reference's own callables (dataset/voc_sample.py:92 `_voc_affine_bboxes`, coco_sample.py:97
    `_coco_affine_bboxes`, visdrone_sample.py:84 `_visdrone_affine_bboxes`, recognised by name), or the strings
    'voc' / 'coco' / 'visdrone'.  The affine parameters are computed inside the kernel, so a callable that is none
    of the three cannot be honoured and raises ValueError.
    NOTE: the reference mutates batch_pred_bbox in place (base_sample.py:124-136); this kernel
    leaves the input untouched (no caller reuses it)."""
    return _ops.recover(batch_pred_bbox, input_size, batch_original_size, affine_kind(affine_func))


_AFFINE_NAMES = {"_voc_affine_bboxes": "voc", "_coco_affine_bboxes": "coco", "_visdrone_affine_bboxes": "visdrone"}


def affine_kind(affine_func) -> str:
    kind = getattr(affine_func, "pq_kind", None)
    if kind is None and isinstance(affine_func, str):
        kind = affine_func
    if kind is None and callable(affine_func):
        kind = _AFFINE_NAMES.get(getattr(affine_func, "__name__", ""))
    if kind not in ("voc", "coco", "visdrone"):
        raise ValueError("affine_func must be one of _voc/_coco/_visdrone_affine_bboxes (the reference's or "
                         "pqdet_b200.base_sample's), got %r" % (affine_func,))
    return kind


def _tag(kind):
    def affine(input_size, batch_original_size):
        raise NotImplementedError("the affine parameters are computed inside the recover kernel")
    affine.pq_kind = kind
    affine.__name__ = "_%s_affine_bboxes" % kind
    return affine


_voc_affine_bboxes = _tag("voc")
_coco_affine_bboxes = _tag("coco")
_visdrone_affine_bboxes = _tag("visdrone")


def recover_bboxes_prediction_voc(batch_pred_bbox, input_size, batch_original_size):
    return recover_bboxes_prediction(batch_pred_bbox, input_size, batch_original_size, _voc_affine_bboxes)


def recover_bboxes_prediction_coco(batch_pred_bbox, input_size, batch_original_size):
    return recover_bboxes_prediction(batch_pred_bbox, input_size, batch_original_size, _coco_affine_bboxes)


def recover_bboxes_prediction_visdrone(batch_pred_bbox, input_size, batch_original_size):
    return recover_bboxes_prediction(batch_pred_bbox, input_size, batch_original_size, _visdrone_affine_bboxes)


# dataset/__init__.py:17-21
RECOVER_BBOXES_REGISTER = {
    'voc': recover_bboxes_prediction_voc,
    'visdrone': recover_bboxes_prediction_visdrone,
    'coco': recover_bboxes_prediction_coco,
}
