#!/usr/bin/env python
"""A/B timing of the fused decode+NMS kernel for differently compiled libraries:
   PQDET_B200_LIB=build/var/lib_x.so python profiles/tools/fused_ab.py [capacity]
Headline workload (1024 VOC-512 images, sparse profile), 2 input sets rotated, CUDA events over 40 launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pqdet_b200 import _ops, synth  # noqa: E402

cap = sys.argv[1] if len(sys.argv) > 1 else "compact"
B, C, size = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024), 20, 512
dev = torch.device("cuda", 0)
if os.environ.get("PQ_L2_FETCH"):
    # cudaLimitMaxL2FetchGranularity (0x05): a hint for the L2's DRAM fetch size (default 64 bytes)
    import ctypes
    torch.cuda.init()
    torch.zeros(1, device=dev)
    rt = None
    for line in open("/proc/self/maps"):
        if "libcudart" in line:
            rt = ctypes.CDLL(line.split()[-1]); break
    val = ctypes.c_size_t(0)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["PQ_L2_FETCH"])))
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print("cudaLimitMaxL2FetchGranularity: rc %d, now %d" % (rc, val.value))
orig = torch.tensor([float(size), float(size)], device=dev)
sets = []
for i in range(2):
    hs = synth.make_heads(B, C, size, "sparse", seed=i, device=dev)
    sets.append((hs,) + _ops.make_heads(hs, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda"))
out = _ops.alloc_fused_outputs(B, 2048, False, dev)
for i in range(6):
    _ops.decode_nms_fused(sets[i % 2][1], sets[i % 2][2], 2048, False, out=out, capacity=cap)
torch.cuda.synchronize()
ts = []
for i in range(40):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    _ops.decode_nms_fused(sets[i % 2][1], sets[i % 2][2], 2048, False, out=out, capacity=cap)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(40):
    _ops.decode_nms_fused(sets[i % 2][1], sets[i % 2][2], 2048, False, out=out, capacity=cap)
e.record()
torch.cuda.synchronize()
b2b = s.elapsed_time(e) * 1e3 / 40
meta = out[2][:3 * B].view(3, B).cpu()
print("%-28s %s: isolated median %.1f us min %.1f us | back to back %.1f us  (kept %d, overflow %d)" % (
    os.path.basename(os.environ.get("PQDET_B200_LIB", "in-tree")), cap, float(np.median(ts)), float(np.min(ts)), b2b,
    int(meta[0].sum()), int((meta[2] != 0).sum())))
