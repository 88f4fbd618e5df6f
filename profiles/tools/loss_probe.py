#!/usr/bin/env python
"""Kernel-only timing of the multi-level loss launch for the BASELINE configs (CUDA events around the raw
C-ABI call, L2 flushed between iterations).  Used to separate kernel time from autograd/graph glue."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops, synth  # noqa: E402
from pqdet_b200.train_dataset import LabelAssigner  # noqa: E402

VIS = [(9, 13), (25, 17), (16, 31), (47, 29), (32, 51), (83, 48), (61, 91), (131, 99), (210, 189)]
CASES = {"B": (16, 20, 512, 1, 12, "l1", None), "C": (64, 10, 608, 20, 200, "l1", VIS),
         "D": (16, 80, 608, 2, 38, "giou", None), "B64": (64, 20, 512, 1, 12, "l1", None)}
STRIDES = (32, 16, 8)


def main():
    dev = torch.device("cuda")
    flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    for name in (sys.argv[1:] or list(CASES)):
        B, C, size, lo, hi, kind, anc = CASES[name]
        gts = synth.make_gt(B, C, size, lo, hi, seed=0)
        la = LabelAssigner(C, device=dev) if anc is None else LabelAssigner(C, anchors=anc, device=dev)
        out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
        t = la.create_label_batch(gts, out_sizes)
        idx = {8: 0, 16: 1, 32: 2}
        labels = [t[idx[s]] for s in STRIDES]
        gl = [t[3 + idx[s]] for s in STRIDES]
        raws = synth.make_train_heads(B, C, size, seed=0, device=dev)

        def call():
            return _ops.loss_levels(raws, labels, gl, C, STRIDES, kind, 0.5, 0.05, True)
        for _ in range(3):
            call()
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); call(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        sp = la.create_sparse_batch(gts, out_sizes)
        owners = [sp.owner[idx[s]] for s in STRIDES]
        sgl = [sp.bboxes[idx[s]] for s in STRIDES]

        def call_sparse():
            return _ops.loss_levels_sparse(raws, owners, sp.gt, sgl, C, STRIDES, kind, 0.5, 0.05, True)
        for _ in range(3):
            call_sparse()
        tsp = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); call_sparse(); e.record(); torch.cuda.synchronize()
            tsp.append(s.elapsed_time(e))
        print("%s: sparse targets        launch %.1f us" % (name, float(np.median(tsp)) * 1e3))
        cells = sum((size // s) ** 2 for s in STRIDES)
        alg = (2 * 12 * (5 + C) + 12 * (6 + C)) * cells * B
        ms = float(np.median(ts))
        print("%s: B=%d C=%d %d  loss_levels launch %.1f us  (alg %.1f MB -> %.0f GB/s)" % (name, B, C, size, ms * 1e3, alg / 1e6, alg / ms / 1e6))


if __name__ == "__main__":
    main()
