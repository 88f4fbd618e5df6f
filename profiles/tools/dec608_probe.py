#!/usr/bin/env python
"""Materialised decode of all levels (pqdet_decode_levels) on 608 x 608 shapes vs VOC-512, launches queued back to back;
PQDET_DECODE_GENERAL=1 = the general kernel."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pqdet_b200 import _ops  # noqa: E402

def ev(fn, reps=6, inner=6):
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

for C, size, nB in ((10, 608, 64), (10, 608, 256), (80, 608, 64), (20, 512, 256)):
    raws = [torch.randn((nB, 3 * (5 + C), size // s, size // s), device="cuda") for s in (32, 16, 8)]
    nbytes = 2 * sum(r.numel() for r in raws) * 4
    t = ev(lambda: _ops.decode_levels(raws, C, (32, 16, 8)))
    os.environ["PQDET_DECODE_GENERAL"] = "1"
    tg = ev(lambda: _ops.decode_levels(raws, C, (32, 16, 8)))
    del os.environ["PQDET_DECODE_GENERAL"]
    print("C=%d %dx%d bs=%d: decode_levels %.0f us = %.2f TB/s (%.2f of HBM peak); general kernel %.0f us" % (
        C, size, size, nB, t * 1e3, nbytes / t / 1e9, nbytes / t / 1e6 / 6499.0, tg * 1e3))
    del raws
