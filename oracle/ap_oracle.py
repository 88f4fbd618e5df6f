"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (pqdet_b200/).

CPU restatement of the evaluator's statistics (eleflea/PQDet eval/evaluator.py):
  init_statics :31-36, add_detections :159-162 (+ tools.PriorityQueue, tools.py:654-679: ordered by
  (-score, insertion index)), add_labels :164-175, AP :64-139, calculate_ap_by_recall_precision :141-157.
Plain Python loops + numpy, arithmetic and dtypes exactly as numpy evaluates the reference's expressions
(numpy >= 2 promotion rules: the installed numpy is the one the reference runs on here).
Pinned against the live reference in tests/test_oracle_vs_reference.py and against tests/golden/ap.npz.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

AP_IOU_THRESHOLDS = np.linspace(0.5, 0.95, 10)          # eval/evaluator.py:13


class ApOracle:
    def __init__(self, num_classes: int):
        self.num_classes = num_classes
        self.dets = defaultdict(list)                    # class -> [(-score, insertion index, file, bbox)]
        self._index = defaultdict(int)
        self.labels = defaultdict(dict)                  # file -> class -> [bboxes, seen, difficult]
        self.gt_count = defaultdict(int)

    def add_detections(self, file_name, bboxes):
        for bbox in bboxes:                              # :161-162; key = -bbox[4], ties by push order
            c = int(bbox[-1])
            self.dets[c].append((-bbox[4], self._index[c], file_name, bbox))
            self._index[c] += 1

    def add_labels(self, file_name, bboxes, diffs):
        classes = bboxes[:, -1].astype(int)
        for c in set(classes):                           # :166-175
            sel = classes == c
            b = bboxes[sel][:, :4]
            d = diffs[sel].astype(bool)
            perm = np.argsort(d)
            b, d = b[perm], d[perm]
            self.labels[file_name][int(c)] = [b, np.zeros((len(AP_IOU_THRESHOLDS), len(b)), bool), d]
            self.gt_count[int(c)] += np.sum(~d)

    def match(self):
        """-> {class: (tp, fp)} float64 arrays (10, n_det) in the order the reference visits the detections."""
        res = {}
        for c, lst in self.dets.items():
            lst = sorted(lst, key=lambda e: (e[0], e[1]))
            tp = np.zeros((len(AP_IOU_THRESHOLDS), len(lst)))
            fp = np.zeros((len(AP_IOU_THRESHOLDS), len(lst)))
            for di, (_, _, f, bbox) in enumerate(lst):
                label = self.labels[f].get(c)
                if label is None:
                    fp[:, di] = 1
                    continue
                BBGT, seen, difficult = label
                bb = bbox[:4]
                ixmin = np.maximum(BBGT[:, 0], bb[0]); iymin = np.maximum(BBGT[:, 1], bb[1])
                ixmax = np.minimum(BBGT[:, 2], bb[2]); iymax = np.minimum(BBGT[:, 3], bb[3])
                iw = np.maximum(ixmax - ixmin + 1., 0.); ih = np.maximum(iymax - iymin + 1., 0.)
                inters = iw * ih
                uni = ((bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) +
                       (BBGT[:, 2] - BBGT[:, 0] + 1.) * (BBGT[:, 3] - BBGT[:, 1] + 1.) - inters)
                overlaps = inters / uni
                for ti, thr in enumerate(AP_IOU_THRESHOLDS):
                    pick, pick_iou = -1, min(thr, 1 - 1e-10)
                    for mi, miou in enumerate(overlaps):
                        if seen[ti, mi]:
                            continue
                        if pick > -1 and not difficult[pick] and difficult[mi]:
                            break
                        if miou < pick_iou:
                            continue
                        pick, pick_iou = mi, miou
                    if difficult[pick]:                  # pick == -1 reads the LAST flag, like the reference
                        continue
                    if pick == -1 or seen[ti, pick]:
                        fp[ti, di] = 1
                        continue
                    tp[ti, di] = 1
                    seen[ti, pick] = True
            res[c] = (tp, fp)
        return res

    @staticmethod
    def ap_from_tp_fp(tp, fp, gt_count):
        fp = np.cumsum(fp, axis=1); tp = np.cumsum(tp, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            rec = tp / gt_count
        prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        mrecs = np.pad(rec, ((0, 0), (1, 1)), constant_values=(0., 1.))
        mpres = np.pad(prec, ((0, 0), (1, 1)), constant_values=0.).tolist()
        n = len(mpres[0])
        for m in mpres:
            for i in range(n - 1, 0, -1):
                if m[i] > m[i - 1]:
                    m[i - 1] = m[i]
        mpres = np.array(mpres)
        return np.sum(np.diff(mrecs) * mpres[:, 1:], axis=1)

    def AP(self):
        raw = np.zeros((self.num_classes, len(AP_IOU_THRESHOLDS)))
        for c, (tp, fp) in self.match().items():
            raw[c] = self.ap_from_tp_fp(tp, fp, self.gt_count[c])
        return raw
