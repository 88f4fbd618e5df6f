#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused decode+NMS kernel (thread 0 of every CTA, clock64 at the phase barriers).
Needs a library built with -DPQ_PHASE_TIMING:  PQDET_B200_LIB=build/var/lib_phase.so python profiles/tools/fused_phases.py [capacity] [B]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pqdet_b200 import _lib, _ops, synth  # noqa: E402

cap = sys.argv[1] if len(sys.argv) > 1 else "compact"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
C, size = 20, 512
dev = torch.device("cuda", 0)
orig = torch.tensor([float(size), float(size)], device=dev)
hs = synth.make_heads(B, C, size, "sparse", seed=0, device=dev)
h, keep = _ops.make_heads(hs, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
out = _ops.alloc_fused_outputs(B, 2048, False, dev)
lib = _lib.load()
fn = lib.pqdet_debug_phase_cycles
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    _ops.decode_nms_fused(h, keep, 2048, False, out=out, capacity=cap)
torch.cuda.synchronize()
fn(buf, 1)
N = 10
for _ in range(N):
    _ops.decode_nms_fused(h, keep, 2048, False, out=out, capacity=cap)
torch.cuda.synchronize()
fn(buf, 1)
names = ["pull image", "1 scan", "2 prefix", "2b slots", "3 fetch+score", "4 offset/segments", "5 class lists", "6 NMS",
         "7 compact kept", "7b rank+write"]
tot = sum(buf[i] for i in range(10))
print("capacity=%s B=%d: mean cycles per image %.0f (%.1f us at 1.965 GHz)" % (cap, B, tot / (N * B), tot / (N * B) / 1965.0))
for i, n in enumerate(names):
    print("  %-20s %8.0f cycles  %5.1f%%" % (n, buf[i] / (N * B), 100.0 * buf[i] / tot))
