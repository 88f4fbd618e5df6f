#!/usr/bin/env python
"""Materialised decode (pqdet_decode_levels) per level subset and channel count: where the fraction of HBM peak goes.
PQDET_DECODE_STAGES caps the TMA kernel's input ring (A/B)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops
def ev(fn, reps=6, inner=6):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner): fn()
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / inner)
    return float(np.median(ts))
cfgs = ((10, 608, 256), (10, 608, 64), (20, 512, 256), (20, 608, 256))
for C, size, nB in cfgs:
    raws = [torch.randn((nB, 3 * (5 + C), size // s, size // s), device="cuda") for s in (32, 16, 8)]
    for name, sel, st in (("all", raws, (32, 16, 8)), ("s16+s8", raws[1:], (16, 8))):
        nbytes = 2 * sum(r.numel() for r in sel) * 4
        t = ev(lambda: _ops.decode_levels(sel, C, st))
        print("stages<=%s C=%d %d bs=%d %-9s %.0f us = %.2f of HBM peak" % (os.environ.get("PQDET_DECODE_STAGES", "8"), C, size, nB, name, t * 1e3, nbytes / t / 1e6 / 6499.0))
