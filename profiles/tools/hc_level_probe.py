#!/usr/bin/env python
"""Head conv + decode per level (pqdet_head_conv_decode): persistent kernel vs PQDET_HEADCONV_GENERAL=1.
   python profiles/tools/hc_level_probe.py [C] [size] [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops
def ev(fn, reps=5, inner=4):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner): fn()
        e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) / inner)
    return float(np.median(ts))
C = int(sys.argv[1]) if len(sys.argv) > 1 else 80
size = int(sys.argv[2]) if len(sys.argv) > 2 else 608
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ch = 3 * (5 + C)
for cin, s in zip((352, 176, 80), (32, 16, 8)):
    f = torch.randn((B, cin, size // s, size // s), device="cuda")
    w = torch.randn((ch, cin), device="cuda") / cin ** 0.5
    b = torch.randn((ch,), device="cuda") * 0.1
    nbytes = f.numel() * 4 + B * (size // s) ** 2 * ch * 4
    t = ev(lambda: _ops.head_conv_decode(f, w, b, C, float(s)))
    os.environ["PQDET_HEADCONV_GENERAL"] = "1"
    tg = ev(lambda: _ops.head_conv_decode(f, w, b, C, float(s)))
    del os.environ["PQDET_HEADCONV_GENERAL"]
    print("C=%d %d bs=%d level %dx%d Cin=%d: %.0f us = %.2f of HBM peak | general kernel %.0f us = %.2f" % (
        C, size, B, size // s, size // s, cin, t * 1e3, nbytes / t / 1e6 / 6499.0, tg * 1e3, nbytes / tg / 1e6 / 6499.0))
