#!/usr/bin/env python
"""Kernel time of the materialising (drop-in) eval route: Decode x 3 levels -> (B,N,5+C), recover -> (B,N,4+C),
batched torch_nms (pqdet_nms_fused).  CUDA events, inputs larger than L2 (B=256: 413 MB of heads)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops, base_sample, synth, tools  # noqa: E402
from pqdet_b200.interpreter import DetectionHead  # noqa: E402

STRIDES = (32, 16, 8)


def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


def main():
    C, size = 20, 512
    dev = torch.device("cuda")
    for B in (64, 256):
        heads = synth.make_heads(B, C, size, "sparse", seed=0, device=dev)
        head = DetectionHead([dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in STRIDES])
        orig = torch.tensor([[float(size), float(size)]], device=dev).repeat(B, 1)
        R = sum(h.numel() for h in heads) * 4
        with torch.no_grad():
            pred = head(heads)
            t_dec = ev(lambda: head(heads))
            rec = base_sample.recover_bboxes_prediction_voc(pred, (size, size), orig)
            t_rec = ev(lambda: base_sample.recover_bboxes_prediction_voc(pred, (size, size), orig))
            out = _ops.alloc_fused_outputs(B, 2048, False, dev)
            t_nms = ev(lambda: _ops.nms_fused(rec, 0.1, 0.45, "auto_cuda", "tv_cuda", 2048, False, out=out))
            t_cp = ev(lambda: pred.clone())
        print("B=%d: torch clone of the decoded tensor (%.0f MB read + write): %.1f us = %.0f GB/s" % (
            B, pred.numel() * 4 / 1e6, t_cp * 1e3, 2 * pred.numel() * 4 / t_cp / 1e6))
        print("B=%d: decode %.1f us (%.0f GB/s of 2R), recover %.1f us (%.0f GB/s), nms_fused %.1f us (%.0f GB/s read) -> %.0f img/s kernel time"
              % (B, t_dec * 1e3, 2 * R / t_dec / 1e6, t_rec * 1e3, (pred.numel() + rec.numel()) * 4 / t_rec / 1e6,
                 t_nms * 1e3, rec.numel() * 4 / t_nms / 1e6, B / ((t_dec + t_rec + t_nms) * 1e-3)))


if __name__ == "__main__":
    main()
