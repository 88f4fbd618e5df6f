#!/usr/bin/env python
"""1x1 head convolution + decode.  Fused tcgen05 kernel (pqdet_head_conv_decode, three levels) against
PyTorch's own convolution (cuDNN/cuBLAS, TF32 allowed - its default) followed by our single-launch decode.
regnetx-600m-fpn shapes: Cin 352 / 176 / 80 at strides 32 / 16 / 8, VOC C=20, 512x512."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pqdet_b200 import _ops  # noqa: E402
from pqdet_b200.interpreter import DetectionHead  # noqa: E402

STRIDES, CINS = (32, 16, 8), (352, 176, 80)


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
    torch.backends.cudnn.allow_tf32 = True
    head = DetectionHead([dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in STRIDES])
    for B in (64, 256):
        feats = [torch.randn((B, c, size // s, size // s), device=dev) for c, s in zip(CINS, STRIDES)]
        ws = [torch.randn((75, c, 1, 1), device=dev) * 0.03 for c in CINS]
        bs = [torch.randn((75,), device=dev) * 0.1 for _ in CINS]
        with torch.no_grad():
            t_conv = ev(lambda: [torch.nn.functional.conv2d(f, w, b) for f, w, b in zip(feats, ws, bs)])
            raws = [torch.nn.functional.conv2d(f, w, b) for f, w, b in zip(feats, ws, bs)]
            t_dec = ev(lambda: head(raws))
            t_fused = ev(lambda: head.forward_from_features(feats, ws, bs))
            per_level = [ev(lambda f=f, w=w, b=b, s=s: _ops.head_conv_decode(f, w, b, C, s)) for f, w, b, s in zip(feats, ws, bs, STRIDES)]
            os.environ["PQDET_HEADCONV_GENERAL"] = "1"
            t_general = ev(lambda: head.forward_from_features(feats, ws, bs))
            del os.environ["PQDET_HEADCONV_GENERAL"]
        x_bytes = sum(f.numel() for f in feats) * 4
        o_bytes = sum(r.numel() for r in raws) * 4
        print("B=%d: torch conv2d x3 %.1f us + decode %.1f us = %.1f us | fused tcgen05 conv+decode %.1f us "
              "(%.0f GB/s of X read + decoded write; %.2fx)" % (B, t_conv * 1e3, t_dec * 1e3, (t_conv + t_dec) * 1e3,
              t_fused * 1e3, (x_bytes + o_bytes) / t_fused / 1e6, (t_conv + t_dec) / t_fused))
        print("      persistent kernel per level (stride 32/16/8): %s us; general kernel (cp.async staging) all levels %.1f us"
              % (" / ".join("%.1f" % (t * 1e3) for t in per_level), t_general * 1e3))


if __name__ == "__main__":
    main()
