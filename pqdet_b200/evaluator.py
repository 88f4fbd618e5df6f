"""Drop-in for the statistics half of eval/evaluator.py:
Evaluator.init_statics / add_detections / add_labels / AP (:31-36, 64-183) with the same results.

The reference's AP() walks every detection of the data set in a Python loop (10 IoU thresholds x the image's
ground truth per detection).  Here the host keeps the bookkeeping - detections ordered per class by
(-score, insertion index) like tools.PriorityQueue, labels grouped per (image, class) with numpy's own argsort -
and the final cumsum / precision / recall / AP arithmetic, which the reference already evaluates vectorised in
numpy; the matching itself runs as one kernel launch (csrc/ap.cu, pqdet_ap_match).

Promotion rules: parity is pinned to NumPy >= 2 (NEP 50; the goldens were generated with numpy 2.3 plus the np.bool /
np.float shims of oracle/ref_harness.py): a float32 detection box keeps the detection's own area
(bb2-bb0+1)*(bb3-bb1+1) in float32.  The unmodified reference needs numpy < 1.24 (it uses np.bool / np.float), where
float32 scalar + Python float promotes to float64; under those legacy rules `uni` is rounded differently, and an IoU
that sits exactly on one of the ten thresholds can flip between tp and fp.  The bit-identical (C,10) table claim holds
for the NumPy 2 semantics the tests run under.
"""
from __future__ import annotations

import ctypes
from collections import namedtuple
from typing import Sequence

import numpy as np
import torch

from . import _lib

AP_IOU_THRESHOLDS = np.linspace(0.5, 0.95, 10)                     # eval/evaluator.py:13
AP = namedtuple('AP', ['mAPs', 'APs', 'AP', 'raw', 'class_names', 'iou_thresholds'])   # tools.py:37


def calculate_ap_by_recall_precision(recs: np.ndarray, precs: np.ndarray) -> np.ndarray:
    """eval/evaluator.py:141-157 (right-to-left running maximum of the precision, then the area)."""
    mrecs = np.pad(recs, ((0, 0), (1, 1)), constant_values=(0., 1.))
    mpres = np.pad(precs, ((0, 0), (1, 1)), constant_values=0.)
    # same values as the reference's right-to-left loop; contiguous, so that np.sum adds in the same (pairwise) order
    mpres = np.ascontiguousarray(np.maximum.accumulate(mpres[:, ::-1], axis=1)[:, ::-1])
    return np.sum(np.diff(mrecs) * mpres[:, 1:], axis=1)


class DetectionAccumulator:
    """evaluator = DetectionAccumulator(class_names); add_detections / add_labels per image as the reference's
    Evaluator.evaluate does (:53-61); AP() returns tools.AP and resets the statistics (:138)."""

    def __init__(self, class_names: Sequence[str], device="cuda"):
        self._classes = list(class_names)
        self.device = torch.device(device)
        self.init_statics()

    def init_statics(self):
        self.detections_count = 0
        self._det_rows, self._det_file = [], []
        self._files = {}
        self._labels = {}                                          # (file id, class) -> (bboxes, difficult)
        self.gt_count = {}

    def _fid(self, file_name) -> int:
        return self._files.setdefault(file_name, len(self._files))

    def add_detections(self, file_name, bboxes):
        """bboxes (K,6) [x1,y1,x2,y2,score,class] as tools.torch_nms returns them (shape (0,) when empty)."""
        if isinstance(bboxes, torch.Tensor):
            bboxes = bboxes.detach().cpu().numpy()
        bboxes = np.asarray(bboxes)
        if bboxes.size == 0:
            return
        bboxes = bboxes.reshape(-1, 6).astype(np.float32, copy=False)
        self.detections_count += len(bboxes)
        self._det_rows.append(bboxes)
        self._det_file.append(np.full((len(bboxes),), self._fid(file_name), np.int64))

    def add_detections_batch(self, file_names, dets):
        """dets: fused.Detections / fused.HostDetections of a whole batch - one device->host copy, then views."""
        for f, rows in zip(file_names, dets.to_numpy_list()):
            self.add_detections(f, rows)

    def add_labels(self, file_name, bboxes: np.ndarray, diffs: np.ndarray):
        """bboxes (n,5) [x1,y1,x2,y2,class], diffs (n,) - eval/evaluator.py:164-175."""
        bboxes, diffs = np.asarray(bboxes), np.asarray(diffs)
        fid = self._fid(file_name)
        classes = bboxes[:, -1].astype(int)
        for class_index in set(classes):
            sel = classes == class_index
            b = bboxes[sel][:, :4]
            d = diffs[sel].astype(bool)
            perm = np.argsort(d)                                   # the reference's own (numpy) ordering
            self._labels[(fid, int(class_index))] = (b[perm], d[perm])
            self.gt_count[int(class_index)] = self.gt_count.get(int(class_index), 0) + np.sum(~d[perm])

    # ------------------------------------------------------------------------------------------------
    def _match(self):
        """-> cls (D,), tp, fp (T,D) uint8 in class-sorted detection order."""
        det = np.concatenate(self._det_rows, axis=0)
        fid = np.concatenate(self._det_file, axis=0)
        D = len(det)
        cls = det[:, 5].astype(np.int64)                           # int(bbox[-1]) truncates like astype
        # tools.PriorityQueue: heap of (-score, push index): ascending -score, ties by insertion order
        order = np.lexsort((np.arange(D), -det[:, 4], cls))
        det, fid, cls = det[order], fid[order], cls[order]
        keys = list(self._labels.keys())
        gid = {k: i for i, k in enumerate(keys)}
        grp = np.fromiter((gid.get((int(f), int(c)), -1) for f, c in zip(fid, cls)), np.int64, D)
        T = len(AP_IOU_THRESHOLDS)
        tp = np.zeros((T, D), np.uint8)
        fp = np.zeros((T, D), np.uint8)
        fp[:, grp < 0] = 1                                         # :72-75: no label of that class in the image
        G = len(keys)
        if G and bool((grp >= 0).any()):
            dts = {np.dtype(self._labels[k][0].dtype) for k in keys}
            if dts == {np.dtype(np.float32)}:
                gdt = np.float32                                   # numpy keeps float32 throughout
            elif np.dtype(np.float32) in dts or np.dtype(np.float16) in dts:
                raise NotImplementedError("ground truth must be all float32, or float64 / integer: numpy would "
                                          "use a different precision per image for a mix (%s)" % sorted(map(str, dts)))
            else:
                gdt = np.float64                                   # float64 and integer labels: float64 arithmetic
            gt_box = np.concatenate([self._labels[k][0].astype(gdt) for k in keys], axis=0)
            gt_diff = np.concatenate([self._labels[k][1] for k in keys]).astype(np.uint8)
            gt_off = np.zeros((G + 1,), np.int64)
            gt_off[1:] = np.cumsum([len(self._labels[k][0]) for k in keys])
            have = np.nonzero(grp >= 0)[0]
            by_grp = have[np.argsort(grp[have], kind="stable")]    # stable: keeps class-rank order inside a group
            grp_det = by_grp.astype(np.int32)
            grp_det_off = np.zeros((G + 1,), np.int64)
            grp_det_off[1:] = np.cumsum(np.bincount(grp[have], minlength=G))
            dev = self.device

            def up(a):
                return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            t_box, t_gd, t_gdo = up(det[:, :4]), up(grp_det), up(grp_det_off)
            t_gt, t_diff, t_off, t_thr = up(gt_box), up(gt_diff), up(gt_off), up(AP_IOU_THRESHOLDS)
            sum_gt = int(gt_off[-1])
            t_seen = torch.zeros((T, sum_gt), dtype=torch.uint8, device=dev)
            t_tp = torch.zeros((T, D), dtype=torch.uint8, device=dev)
            t_fp = torch.zeros((T, D), dtype=torch.uint8, device=dev)
            p = lambda x: ctypes.c_void_p(x.data_ptr())
            dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
            _lib.check(_lib.load().pqdet_ap_match(
                p(t_box), D, p(t_gd), p(t_gdo), p(t_gt), 1 if gdt is np.float64 else 0, p(t_diff), p(t_off), sum_gt, G,
                p(t_thr), T, p(t_seen), p(t_tp), p(t_fp), dev_index,
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "pqdet_ap_match")
            tp = t_tp.cpu().numpy()
            fp |= t_fp.cpu().numpy()
        return cls, tp, fp

    def AP(self) -> AP:
        if not self._det_rows:
            # the reference leaves `metrics` unassigned when there is no detection at all (:136, 139)
            raise UnboundLocalError("cannot access local variable 'metrics' where it is not associated with a value")
        cls, tp, fp = self._match()
        raw = np.zeros((len(self._classes), len(AP_IOU_THRESHOLDS)))
        for c in np.unique(cls):
            m = cls == c
            # C-contiguous (T, n) like the reference's arrays: np.sum's pairwise order depends on the memory layout
            fpc = np.cumsum(np.ascontiguousarray(fp[:, m], dtype=np.float64), axis=1)
            tpc = np.cumsum(np.ascontiguousarray(tp[:, m], dtype=np.float64), axis=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                rec = tpc / self.gt_count.get(int(c), 0)
            prec = tpc / np.maximum(tpc + fpc, np.finfo(np.float64).eps)
            raw[int(c)] = calculate_ap_by_recall_precision(rec, prec)
        APs = np.mean(raw, axis=1)
        mAPs = np.mean(raw, axis=0)
        metrics = AP(mAPs, APs, np.mean(mAPs), raw, self._classes, AP_IOU_THRESHOLDS)
        self.init_statics()
        return metrics
