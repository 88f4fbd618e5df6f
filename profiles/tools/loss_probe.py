#!/usr/bin/env python
"""The one-launch loss step (pqdet_loss_levels / _sparse) on BASELINE configs B, C, D: CUDA-event time of the eager
call with an L2 flush between launches, for A/B runs (PQDET_B200_LIB) and as the ncu target:
    python profiles/tools/loss_probe.py [B|C|D] [dense|sparse] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pqdet_b200 import config as pqcfg, synth  # noqa: E402
from pqdet_b200.interpreter import DetectionHead  # noqa: E402
from pqdet_b200.train_dataset import LabelAssigner, assign_labels, assign_sparse, pack_gt  # noqa: E402

VIS = [(9, 13), (25, 17), (16, 31), (47, 29), (32, 51), (83, 48), (61, 91), (131, 99), (210, 189)]
CFG = {"B": (16, 20, 512, 1, 12, "l1", None), "C": (64, 10, 608, 20, 200, "l1", "vis"),
       "D": (16, 80, 608, 2, 38, "giou", None)}
which = sys.argv[1] if len(sys.argv) > 1 else "C"
form = sys.argv[2] if len(sys.argv) > 2 else "dense"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
B, C, size, lo, hi, kind, anchors = CFG[which]
dev = torch.device("cuda", 0)
pqcfg.nan_check = "off"
gts = synth.make_gt(B, C, size, lo, hi, seed=0)
out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
la = LabelAssigner(C, device=dev)
if anchors == "vis":
    la = LabelAssigner(C, anchors=VIS, device=dev)
gt_dev, cnt_dev = pack_gt(gts, dev)
fn = assign_sparse if form == "sparse" else assign_labels
target = fn(gt_dev, cnt_dev, out_sizes, C, la._anchors.tolist(), la._strides.tolist(), 0.3)
raws = synth.make_train_heads(B, C, size, seed=0, device=dev)
opts = [dict(classes=C, stride=s, bbox_loss=kind, ignore_thresh=0.5, l1_loss_gain=0.05) for s in (32, 16, 8)]
head = DetectionHead(opts)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
for _ in range(3):
    head.loss_and_grad(raws, target)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out, grads = head.loss_and_grad(raws, target)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
from pqdet_b200.graphs import GraphedLossStep  # noqa: E402
g = GraphedLossStep(head, [r.requires_grad_(True) for r in raws], target)
tg = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    tg.append(s.elapsed_time(e) * 1e3)
print("%s config %s %s: CUDA-graph replay median %.1f us min %.1f us" % (
    os.path.basename(os.environ.get("PQDET_B200_LIB", "in-tree")), which, form, float(np.median(tg)), float(np.min(tg))))
print("%s config %s %s: loss step median %.1f us min %.1f us, loss %.6f" % (
    os.path.basename(os.environ.get("PQDET_B200_LIB", "in-tree")), which, form, float(np.median(ts)), float(np.min(ts)),
    float(out["loss"])))
