"""Drop-in for the target assignment of dataset/train_dataset.py:13-43, 109-150.

The reference builds dense labels with numpy inside DataLoader workers (2.5-11.5 ms / image) and
then ships ~L bytes per image host->device every step.  Here only the GT boxes (n,6) cross PCIe;
the dense (B,H,W,3,6+C) tensors are produced on the GPU by csrc/assign.cu in the reference's mixed
fp32/fp64 arithmetic, bit-identical to numpy's result.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _ops

DEFAULT_ANCHORS = [(10, 13), (16, 30), (33, 23), (30, 61), (62, 45), (59, 119),
                   (116, 90), (156, 198), (373, 326)]           # config.py:58-59
DEFAULT_STRIDES = (8, 16, 32)                                   # config.py:56


def pack_gt(batch_bboxes: Sequence[np.ndarray], device) -> tuple:
    """list of (n_i, 6) arrays [x1,y1,x2,y2,class,mixw] -> padded (B, n_max, 6) + counts, on device."""
    B = len(batch_bboxes)
    n_max = max([len(b) for b in batch_bboxes] + [1])
    gt = np.zeros((B, n_max, 6), dtype=np.float32)
    cnt = np.zeros((B,), dtype=np.int32)
    for i, b in enumerate(batch_bboxes):
        b = np.asarray(b, dtype=np.float32).reshape(-1, 6)
        gt[i, :len(b)] = b
        cnt[i] = len(b)
    pin = torch.cuda.is_available()
    gt_t, cnt_t = torch.from_numpy(gt), torch.from_numpy(cnt)
    if pin:
        gt_t, cnt_t = gt_t.pin_memory(), cnt_t.pin_memory()
    return gt_t.to(device, non_blocking=True), cnt_t.to(device, non_blocking=True)


def assign_labels(gt: torch.Tensor, gt_count: torch.Tensor, output_sizes, num_classes: int,
                  anchors=DEFAULT_ANCHORS, strides=DEFAULT_STRIDES, anchors_iou_threshold: float = 0.3,
                  trim: bool = True):
    """Batched create_label + collate_batch on the GPU.
    gt (B, n_max, 6) CUDA, gt_count (B).  output_sizes (3,2) = (h,w) per scale (strides ascending).
    -> (label_s, label_m, label_l, sbboxes, mbboxes, lbboxes) with labels (B,H,W,3,6+C) and GT lists
    (B, G_s, 4), G_s = the batch maximum (>= 1) like _pad_arrays (train_dataset.py:21-24).
    trim=False skips the one host sync needed to size G_s and returns the lists at capacity 3*n_max
    (zero rows beyond the true length do not change the loss: IoU with a zero box is 0)."""
    labels, lists, list_len = _ops.assign_labels(gt, gt_count, num_classes, anchors, strides,
                                                 [tuple(int(v) for v in s) for s in output_sizes],
                                                 anchors_iou_threshold)
    if trim:
        mx = list_len.max(dim=0)[0].cpu().tolist()
        lists = [l[:, :max(int(m), 1)].contiguous() for l, m in zip(lists, mx)]
    return (labels[0], labels[1], labels[2], lists[0], lists[1], lists[2])


class SparseTarget:
    """Training target without dense labels: what DetectionHead.forward accepts in
    place of the 6-tuple (label_s, label_m, label_l, sbboxes, mbboxes, lbboxes).
      gt      (B, n_max, 6)  the GT rows [x1,y1,x2,y2,class,mixw] themselves
      owner   3 x (B, 3, H_s, W_s) int32, scales in ascending stride order (8, 16, 32): GT index or -1
      bboxes  3 x (B, G_s, 4) per-scale GT lists (what loss_per_scale's ignore mask consumes)
    The loss kernel rebuilds each responsible cell's label row from gt[owner]; results are bit-identical to the
    dense path, L bytes per image are neither written nor read."""

    def __init__(self, gt, owner, bboxes, num_classes):
        self.gt, self.owner, self.bboxes, self.num_classes = gt, list(owner), list(bboxes), num_classes

    def tensors(self):
        return [self.gt] + self.owner + self.bboxes

    def clone(self):
        return SparseTarget(self.gt.clone(), [o.clone() for o in self.owner], [b.clone() for b in self.bboxes],
                            self.num_classes)

    def copy_(self, other):
        for d, s in zip(self.tensors(), other.tensors()):
            if d.shape != s.shape:
                raise ValueError("target shape changed: %s vs %s (use trim=False for a fixed capacity)"
                                 % (tuple(s.shape), tuple(d.shape)))
            d.copy_(s, non_blocking=True)
        return self


def assign_sparse(gt: torch.Tensor, gt_count: torch.Tensor, output_sizes, num_classes: int,
                  anchors=DEFAULT_ANCHORS, strides=DEFAULT_STRIDES, anchors_iou_threshold: float = 0.3,
                  trim: bool = True) -> SparseTarget:
    """The assignment of create_label + collate_batch without the dense label tensors."""
    owners, lists, list_len = _ops.assign_sparse(gt, gt_count, anchors, strides,
                                                 [tuple(int(v) for v in s) for s in output_sizes],
                                                 anchors_iou_threshold)
    if trim:
        mx = list_len.max(dim=0)[0].cpu().tolist()
        lists = [l[:, :max(int(m), 1)].contiguous() for l, m in zip(lists, mx)]
    return SparseTarget(gt, owners, lists, num_classes)


class LabelAssigner:
    """Holds what TrainDataset.__init__ reads from the config (train_dataset.py:47-56) and exposes
    create_label with the reference's signature."""

    def __init__(self, num_classes: int, anchors=DEFAULT_ANCHORS, strides=DEFAULT_STRIDES,
                 anchors_iou_threshold: float = 0.3, device="cuda"):
        self._num_classes = num_classes
        self._anchors = np.array(anchors, dtype=np.float32)
        self._strides = np.array(strides)
        self._anchors_iou_threshold = anchors_iou_threshold
        self._gt_per_grid = 3
        self.device = torch.device(device)

    def create_label(self, bboxes, output_sizes):
        """train_dataset.py:109-150 for ONE image: bboxes (n,6) -> (label_s, label_m, label_l,
        sbboxes, mbboxes, lbboxes); labels are CUDA tensors (H,W,3,6+C), GT lists are (g,4) tensors
        (g may be 0), so that collate_batch below can pad them."""
        gt, cnt = pack_gt([np.asarray(bboxes, dtype=np.float32).reshape(-1, 6)], self.device)
        labels, lists, list_len = _ops.assign_labels(gt, cnt, self._num_classes, self._anchors.tolist(),
                                                     self._strides.tolist(),
                                                     [tuple(int(v) for v in s) for s in output_sizes],
                                                     self._anchors_iou_threshold)
        n = list_len[0].cpu().tolist()
        return (labels[0][0], labels[1][0], labels[2][0],
                lists[0][0, :n[0]], lists[1][0, :n[1]], lists[2][0, :n[2]])

    def create_label_batch(self, batch_bboxes: List[np.ndarray], output_sizes, trim: bool = True):
        gt, cnt = pack_gt(batch_bboxes, self.device)
        return assign_labels(gt, cnt, output_sizes, self._num_classes, self._anchors.tolist(),
                             self._strides.tolist(), self._anchors_iou_threshold, trim=trim)

    def create_sparse_batch(self, batch_bboxes: List[np.ndarray], output_sizes, trim: bool = True) -> SparseTarget:
        gt, cnt = pack_gt(batch_bboxes, self.device)
        return assign_sparse(gt, cnt, output_sizes, self._num_classes, self._anchors.tolist(),
                             self._strides.tolist(), self._anchors_iou_threshold, trim=trim)


def _pad_arrays(arrays):
    """train_dataset.py:21-24 on tensors: zero-pad (g,4) lists to the batch max (>= 1)."""
    n = max(max(len(a) for a in arrays), 1)
    out = arrays[0].new_zeros((len(arrays), n, 4))
    for i, a in enumerate(arrays):
        if len(a):
            out[i, :len(a)] = a
    return out


def collate_batch(batch):
    """train_dataset.py:26-43 for samples produced by LabelAssigner.create_label:
    each sample = (image, label_s, label_m, label_l, sbboxes, mbboxes, lbboxes)."""
    transposed = list(zip(*batch))
    images_and_labels = [torch.stack([torch.as_tensor(x) for x in samples], 0) for samples in transposed[:4]]
    bboxs = [_pad_arrays(list(samples)) for samples in transposed[4:]]
    return (*images_and_labels, *bboxs)
