#!/usr/bin/env python
"""Dense label assignment (pqdet_assign_labels) and the sparse form on BASELINE configs C and D, L2 flushed between calls;
PQDET_ASSIGN_NO_PDL=1 = the owner pass strictly after the background fill."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import synth
from pqdet_b200.train_dataset import LabelAssigner, assign_labels, assign_sparse, pack_gt
VIS = [(9, 13), (25, 17), (16, 31), (47, 29), (32, 51), (83, 48), (61, 91), (131, 99), (210, 189)]
dev = torch.device("cuda", 0)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
for name, (B, C, size, lo, hi, anchors) in {"C": (64, 10, 608, 20, 200, VIS), "D": (16, 80, 608, 2, 38, None), "B": (16, 20, 512, 1, 12, None)}.items():
    gts = synth.make_gt(B, C, size, lo, hi, seed=0)
    out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
    la = LabelAssigner(C, device=dev) if anchors is None else LabelAssigner(C, anchors=anchors, device=dev)
    gt_dev, cnt_dev = pack_gt(gts, dev)
    L = B * sum((size // s) ** 2 for s in (8, 16, 32)) * 3 * (6 + C) * 4
    for form, fn in (("dense", assign_labels), ("sparse", assign_sparse)):
        ts = []
        for _ in range(12):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(gt_dev, cnt_dev, out_sizes, C, la._anchors.tolist(), la._strides.tolist(), 0.3, trim=False)
            e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        t = float(np.median(ts[2:]))
        print("config %s %-6s assignment: %.1f us%s" % (name, form, t, (" = %.2f of HBM peak on the %d MB of labels" % (L / (t * 1e-6) / 1e9 / 6499.0, L >> 20)) if form == "dense" else ""))
