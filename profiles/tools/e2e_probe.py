#!/usr/bin/env python
"""The host-buffers-in / host-buffers-out call (pqdet_decode_nms_host) on the headline workload: capacity classes,
and the batch split into halves / quarters on separate streams."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops, synth
B, C, size = 1024, 20, 512
dev = torch.device("cuda", 0)
hs = synth.make_heads(B, C, size, "sparse", seed=0, device="cpu")
hh = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in hs]
orig = torch.tensor([float(size), float(size)])
def wall(fn, n=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))
for cap in ("compact", "large"):
    hz, keep = _ops.make_heads_host(hh, (32, 16, 8), C, (size, size), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
    out = _ops.alloc_host_outputs(B, 2048, False, dev)
    t = wall(lambda: _ops.decode_nms_host(hz, keep, 2048, False, dev, out=out, capacity=cap))
    print("%s capacity %-8s one launch: %.2f ms = %.0f k images/s" % (os.path.basename(os.environ.get("PQDET_B200_LIB", "in-tree")), cap, t * 1e3, B / t / 1e3))
for parts in ((2, 4, 8) if not os.environ.get("PQDET_B200_LIB") else ()):
    n = B // parts
    streams = [torch.cuda.Stream() for _ in range(parts)]
    subs = []
    for p in range(parts):
        sub = [t[p * n:(p + 1) * n] for t in hh]
        hz, keep = _ops.make_heads_host(sub, (32, 16, 8), C, (size, size), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
        subs.append((hz, keep, _ops.alloc_host_outputs(n, 2048, False, dev)))
    def run():
        for st, (hz, keep, out) in zip(streams, subs):
            with torch.cuda.stream(st):
                _ops.decode_nms_host(hz, keep, 2048, False, dev, out=out)
    t = wall(run)
    print("%d launches of %d images on %d streams: %.2f ms = %.0f k images/s" % (parts, n, parts, t * 1e3, B / t / 1e3))
