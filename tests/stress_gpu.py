"""Randomised sweep of the eval path on a GPU box (not collected by pytest; a few minutes):

    python tests/stress_gpu.py [n_cases] [seed]

Every case draws a class count, input size, FPN/PAN level order, thresholds, affine kind, original image sizes and a
synthetic profile, then checks the fused kernel, the general path, the batched torch_nms drop-in and the host-buffer
entry point against the oracle (bit-exact rows on our own decoded boxes), exactly like tests/test_gpu_nms.py does for
its fixed cases."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gpu_util import assert_same_detections, cuda, eval_chain_oracle  # noqa: E402
from oracle import pqdet_oracle as po  # noqa: E402
from pqdet_b200 import base_sample, config, fused, synth, tools  # noqa: E402
from pqdet_b200.interpreter import DetectionHead  # noqa: E402


def one_case(rng, i):
    C = int(rng.choice([1, 2, 3, 10, 20, 80]))
    size = int(rng.choice([256, 288, 320, 352, 416, 512]))
    strides = (32, 16, 8) if rng.random() < 0.7 else (8, 16, 32)
    profile = str(rng.choice(["sparse", "sparse", "sparse", "dense"]))
    if profile == "dense" and C > 20:
        profile = "sparse"
    thr = float(rng.choice([0.05, 0.1, 0.25, 0.5]))
    iou = float(rng.choice([0.3, 0.45, 0.5, 0.65]))
    kind = str(rng.choice(["voc", "coco", "visdrone"]))
    sem = str(rng.choice(["cuda", "cpu"]))
    B = int(rng.integers(1, 5))
    orig = np.stack([rng.integers(120, 900, 2).astype(np.float32) for _ in range(B)])
    config.nms_semantics = sem
    heads = synth.make_heads(B, C, size, profile, seed=int(rng.integers(0, 1 << 30)), strides=strides)
    dheads = [h.cuda() for h in heads]
    opts = [dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in strides]
    decoded_t = DetectionHead(opts)(dheads)
    decoded = decoded_t.cpu().numpy()
    want = eval_chain_oracle(None, strides, C, (size, size), orig, kind, thr, iou, sem, decoded=decoded)
    dets = fused.decode_nms(dheads, strides, C, (size, size), cuda(orig), kind, thr, iou, return_index=True)
    gen = fused.decode_nms(dheads, strides, C, (size, size), cuda(orig), kind, thr, iou, return_index=True, strategy="general")
    host = fused.decode_nms_host([h.pin_memory() for h in heads], strides, C, (size, size), torch.from_numpy(orig), kind, thr, iou)
    rec = base_sample.RECOVER_BBOXES_REGISTER[kind](decoded_t, (size, size), cuda(orig))
    drop = tools.batched_torch_nms(rec, thr, iou)
    hrows = host.to_numpy_list()
    for b in range(B):
        w, rows, cls = want[b]
        ncand = int(dets.host_meta()[1, b])
        vanilla = 4 * ncand > (100000 if sem == "cuda" else 4000)
        tag = "case %d img %d (C=%d size=%d %s %s thr=%g iou=%g %s)" % (i, b, C, size, profile, kind, thr, iou, sem)
        assert_same_detections(dets[b].cpu().numpy(), w, ties_unordered=vanilla, what=tag)
        assert torch.equal(gen[b], dets[b]), tag + " general"
        assert np.array_equal(hrows[b], dets[b].cpu().numpy()), tag + " host"
        d = drop[b].cpu().numpy().reshape(-1, 6) if drop[b].numel() else np.zeros((0, 6), np.float32)
        assert_same_detections(d, w, ties_unordered=vanilla, what=tag + " drop-in")
    return sum(int(dets.host_meta()[0, b]) for b in range(B))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    kept = 0
    for i in range(n):
        kept += one_case(rng, i)
    print("stress ok: %d cases, %d detections compared bit-exactly" % (n, kept))


if __name__ == "__main__":
    main()
