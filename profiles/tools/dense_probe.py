#!/usr/bin/env python
"""Where the time goes for BASELINE config C (VisDrone-shaped dense heads, bs=64): fused attempt, the
general path alone (kernel time by CUDA events), and the end-to-end fused.decode_nms wall time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops, fused, synth  # noqa: E402

STRIDES = (32, 16, 8)


def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


def main():
    B, C, size = 64, 10, 608
    dev = torch.device("cuda")
    heads = synth.make_heads(B, C, size, "dense", seed=0, device=dev)
    orig = torch.tensor([480.0, 480.0], device=dev)
    h, keep = _ops.make_heads(heads, STRIDES, C, (size, size), orig, "visdrone", 0.1, 0.45, "auto_cuda", "tv_cuda")
    out = _ops.alloc_fused_outputs(B, 2048, False, dev)
    print("fused attempt (all images overflow): %.1f us" % (1e3 * ev(lambda: _ops.decode_nms_fused(h, keep, 2048, False, out=out))))
    print("general path, all 64 images: %.1f us" % (1e3 * ev(lambda: _ops.nms_general(heads_t=h, keep_alive=keep, n_images=B, max_det=8192, cand_capacity=1 << 21))))
    def full():
        return fused.decode_nms(heads, STRIDES, C, (size, size), orig, "visdrone", 0.1, 0.45)
    for strat in ("fused", "general", "auto"):
        ts = []
        for _ in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d = fused.decode_nms(heads, STRIDES, C, (size, size), orig, "visdrone", 0.1, 0.45, strategy=strat)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e6)
        print("fused.decode_nms strategy=%s wall per call (us): %s" % (strat, " ".join("%.0f" % t for t in ts)))
    print("kept/img %.0f cand/img %.0f" % (float(d.host_meta()[0].float().mean()), float(d.host_meta()[1].float().mean())))


if __name__ == "__main__":
    main()
