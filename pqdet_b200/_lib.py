"""ctypes binding of the C ABI (include/pqdet_b200.h).

There is no CPU fallback: if the shared library is missing or a call is made without a CUDA
tensor, this module raises.  The library is built in-tree by pqdet_b200/build.py.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

from . import build as _build

MAX_LEVELS = 4
MAX_CLASSES = 126

AFFINE = {"voc": 0, "coco": 1, "visdrone": 2}
NMS_MODE = {"auto_cuda": 0, "auto_cpu": 1, "trick": 2, "vanilla": 3}
IOU_ROUND = {"tv_cuda": 0, "tv_cpu": 1}
BBOX_LOSS = {"l1": 0, "iou": 1, "giou": 2, "diou": 3}
ST_CAND_OVERFLOW, ST_DET_TRUNCATED = 1, 2
# capacity classes of the fused kernels (include/pqdet_b200.h): per-image (hit rows, candidates) kept on chip
CAPACITY = {"compact": 0, "large": 1}
CAPACITY_LIMITS = {"compact": (512, 1280), "large": (1024, 2048)}


PQDET_ERR_UNSUPPORTED = -3      # include/pqdet_b200.h


class PqdetError(RuntimeError):
    pass


class HeadsT(ctypes.Structure):
    _fields_ = [
        ("raw", c_void_p * MAX_LEVELS),
        ("H", c_int * MAX_LEVELS), ("W", c_int * MAX_LEVELS),
        ("stride", c_float * MAX_LEVELS),
        ("n_levels", c_int),
        ("B", c_int), ("A", c_int), ("C", c_int),
        ("affine_kind", c_int),
        ("in_h", c_float), ("in_w", c_float),
        ("orig_hw", c_void_p),
        ("orig_per_image", c_int),
        ("score_threshold", c_double),
        ("iou_threshold", c_double),
        ("nms_mode", c_int), ("iou_round", c_int),
    ]


# name -> (restype, argtypes); every symbol declared in include/pqdet_b200.h
SIGNATURES = {
    "pqdet_version": (c_int, []),
    "pqdet_strerror": (c_char_p, [c_int]),
    "pqdet_decode_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_int64, c_int64, c_int, c_void_p]),
    "pqdet_decode_levels": (c_int, [c_int, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_float),
                                    c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "pqdet_decode_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_int64, c_int64, c_int, c_void_p]),
    "pqdet_recover": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float, c_float,
                              c_void_p, c_int, c_int, c_void_p]),
    "pqdet_decode_nms": (c_int, [POINTER(HeadsT), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int, c_int, c_void_p]),
    "pqdet_decode_nms_host": (c_int, [POINTER(HeadsT), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int, c_int, c_int, c_void_p]),
    "pqdet_decode_nms_gather": (c_int, [POINTER(HeadsT), c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_int,
                                        c_void_p, c_int, c_int, c_int, c_void_p]),
    "pqdet_peer_wait": (c_int, [c_void_p, c_int, ctypes.c_uint32, c_void_p, c_int, c_void_p]),
    "pqdet_peer_publish": (c_int, [c_void_p, c_int, c_float, POINTER(c_void_p), POINTER(c_void_p), c_int, c_int, c_int,
                                   c_int, c_void_p]),
    "pqdet_peer_sum_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pqdet_head_conv_hits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_double,
                                     c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pqdet_records_nms": (c_int, [POINTER(HeadsT), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pqdet_nms_fused": (c_int, [c_void_p, c_int, c_int64, c_int, c_double, c_double, c_int, c_int, c_void_p,
                                c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pqdet_nms_general_workspace": (c_int64, [c_int, c_int64, c_int, c_int64, c_int]),
    "pqdet_nms_general": (c_int, [POINTER(HeadsT), c_void_p, c_int64, c_int, c_int, c_double, c_double,
                                  c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "pqdet_classwise_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_double, c_double, c_double,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pqdet_iou_pairwise": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "pqdet_iou_pairwise_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                       c_int, c_void_p]),
    "pqdet_loss_workspace": (c_int64, [c_int, c_int, c_int, c_int]),
    "pqdet_loss_fwd_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                   c_float, c_float, c_int, c_void_p]),
    "pqdet_loss_scale_grad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_void_p]),
    "pqdet_loss_levels_workspace": (c_int64, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "pqdet_loss_levels": (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                  POINTER(c_float), c_int, c_int, c_int, c_int, c_float, c_float, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pqdet_loss_levels_scale_grad": (c_int, [c_int, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), c_int,
                                             c_int, c_int, c_void_p, c_int, c_void_p]),
    "pqdet_assign_workspace": (c_int64, [c_int, POINTER(c_int), POINTER(c_int)]),
    "pqdet_assign_labels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_float),
                                    POINTER(c_int), POINTER(c_int), POINTER(c_int), c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p, c_int, c_void_p]),
    "pqdet_assign_sparse": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_float), POINTER(c_int),
                                    POINTER(c_int), POINTER(c_int), c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "pqdet_loss_levels_sparse": (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_int,
                                         POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), POINTER(c_int),
                                         POINTER(c_int), POINTER(c_float), c_int, c_int, c_int, c_int, c_float,
                                         c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pqdet_ap_match": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64,
                               c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pqdet_letterbox_normalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(c_float),
                                          POINTER(c_float), c_void_p, c_void_p, c_int, c_void_p]),
    "pqdet_head_conv_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_int, c_float, c_int64, c_int64, c_int, c_void_p]),
    "pqdet_head_conv_decode_levels": (c_int, [c_int, c_void_p, c_void_p, c_void_p, POINTER(c_int), POINTER(c_int),
                                              POINTER(c_int), POINTER(c_float), c_void_p, c_int, c_int, c_int, c_int,
                                              c_void_p]),
}

_LIB = None


def lib_path() -> str:
    """The in-tree library; PQDET_B200_LIB overrides it (A/B runs of differently compiled kernels)."""
    return os.environ.get("PQDET_B200_LIB") or _build.LIB_PATH


def load():
    """Load (never build implicitly on a GPU box: the .so travels with the tree)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PqdetError(
            "pqdet_b200 CUDA extension not built: %s is missing. Run `python -m pqdet_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().pqdet_strerror(int(code)).decode()
        raise PqdetError("%s failed: %s (code %d)" % (what or "pqdet call", msg, code))
