#!/usr/bin/env python
"""Head conv + decode on 608 x 608 levels (BASELINE config C shape: 10 classes, Cin 352/176/80): the persistent
kernel with partial tiles / cooperative stores against round 1's eligibility rule (PQDET_HEADCONV_ALIGNED_ONLY=1: these
levels on the general kernel), launches queued back to back; features -> detections beside it."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pqdet_b200 import fused, synth  # noqa: E402
from pqdet_b200.interpreter import DetectionHead  # noqa: E402

def ev(fn, reps=6, inner=4):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner):
            fn()
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / inner)
    return float(np.median(ts))

PEAK = 6499.0
for C, size, cins, nB in ((10, 608, (352, 176, 80), 64), (10, 608, (352, 176, 80), 256), (20, 512, (352, 176, 80), 256),
                          (80, 608, (352, 176, 80), 64)):
    ch = 3 * (5 + C)
    strides = synth.FPN_STRIDES
    feats = [torch.randn((nB, c, size // s, size // s), device="cuda") for c, s in zip(cins, strides)]
    ws = [torch.randn((ch, c), device="cuda") / c ** 0.5 for c in cins]
    bs = [torch.randn((ch,), device="cuda") * 0.1 for _ in cins]
    for b_ in bs:
        b_[4::(5 + C)] -= 4.6
    head = DetectionHead([dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in strides])
    xb = sum(f.numel() for f in feats) * 4
    ob = nB * sum((size // s) ** 2 for s in strides) * ch * 4
    orig = torch.tensor([float(size), float(size)], device="cuda")
    with torch.no_grad():
        t_new = ev(lambda: head.forward_from_features(feats, ws, bs))
        os.environ["PQDET_HEADCONV_ALIGNED_ONLY"] = "1"
        t_old = ev(lambda: head.forward_from_features(feats, ws, bs))
        del os.environ["PQDET_HEADCONV_ALIGNED_ONLY"]
        t_det = ev(lambda: fused.features_nms(feats, ws, bs, strides, C, (size, size), orig, "coco", 0.1, 0.45), inner=1)
    print("C=%d %dx%d bs=%d: head conv + decode %.0f us = %.2f TB/s (%.2f of HBM peak); round-1 eligibility %.0f us; "
          "features -> detections %.0f us (%.2f TB/s of feature reads)" % (
              C, size, size, nB, t_new * 1e3, (xb + ob) / t_new / 1e9, (xb + ob) / t_new / 1e6 / PEAK, t_old * 1e3,
              t_det * 1e3, xb / t_det / 1e9))
    del feats
