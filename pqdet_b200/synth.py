"""PQ-SYNTH-v1: synthetic YOLO-head outputs and ground-truth boxes.

Bench/test input generator only -- no detection math lives here.  Background logits are
Gaussian noise drawn with a seeded torch generator on the target device; "planted" objects
(drawn with a seeded numpy generator, so they are device independent) overwrite the 3x3 cell
neighbourhood of their centre on the FPN level that matches their size.

Profiles: sparse (2..8 objects / image, VOC/COCO-like), dense (80..300, VisDrone-like),
gauss (N(0,1) everywhere, no planting: candidate-overflow stress).
"""
from __future__ import annotations

import numpy as np
import torch

PROFILES = {"sparse": (2, 8), "coco": (2, 20), "dense": (80, 300), "gauss": None}
FPN_STRIDES = (32, 16, 8)          # cfg order of the FPN models


def draw_objects(rng: np.random.Generator, n: int, size: int, num_classes: int):
    """n boxes [x1,y1,x2,y2] (fp32) + classes, as in Appendix B."""
    wh = np.exp(rng.standard_normal((n, 2)) * 0.9 + 3.6).clip(8, 0.8 * size)
    c = rng.random((n, 2)) * size
    cls = rng.integers(0, num_classes, size=n)
    x1y1 = (c - wh / 2).clip(0, size - 2)
    x2y2 = (c + wh / 2).clip(2, size - 1)
    return np.concatenate([x1y1, x2y2], axis=1).astype(np.float32), cls.astype(np.int64)


def make_heads(batch: int, num_classes: int, size: int, profile: str = "sparse", seed: int = 0,
               device="cpu", strides=FPN_STRIDES, anchors_per_cell: int = 3):
    """-> list of raw head tensors (B, A*(5+C), H, W), one per stride, NCHW contiguous fp32."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    A, ch = anchors_per_cell, 5 + num_classes
    heads = []
    for s in strides:
        h = size // s
        r = torch.empty((batch, A, ch, h, h), dtype=torch.float32, device=dev)
        if profile == "gauss":
            r.normal_(0.0, 1.0, generator=g)
        else:
            r[:, :, 0:4].normal_(0.7, 0.5, generator=g)
            r[:, :, 4].normal_(-6.0, 1.0, generator=g)
            r[:, :, 5:].normal_(-3.0, 1.0, generator=g)
        heads.append(r)
    if profile != "gauss":
        lo, hi = PROFILES[profile]
        rng = np.random.default_rng(seed)
        counts = rng.integers(lo, hi + 1, size=batch)
        per_level = [[] for _ in strides]                    # (flat index, value) pairs
        for b in range(batch):
            n = int(counts[b])
            boxes, cls = draw_objects(rng, n, size, num_classes)
            m = np.maximum(boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1])
            level = np.where(m > 160, 0, np.where(m > 64, 1, 2))
            for k in range(n):
                li = int(level[k]) if len(strides) == 3 else 0
                s = strides[li]
                h = size // s
                x1, y1, x2, y2 = boxes[k]
                cx, cy = int(((x1 + x2) / 2) // s), int(((y1 + y2) / 2) // s)
                ys = np.arange(max(cy - 1, 0), min(cy + 1, h - 1) + 1)
                xs = np.arange(max(cx - 1, 0), min(cx + 1, h - 1) + 1)
                yy, xx = np.meshgrid(ys, xs, indexing="ij")
                yy, xx = yy.ravel(), xx.ravel()
                gx, gy = (xx + 0.5) * s, (yy + 0.5) * s
                dist = np.stack([gx - x1, gy - y1, x2 - gx, y2 - gy], axis=0).clip(1.0, None)
                base = np.log(dist / s)                                    # (4, ncell)
                for a in range(A):
                    vals = np.concatenate([
                        base + rng.standard_normal(base.shape) * 0.05,
                        rng.standard_normal((1, yy.size)) + 2.5,
                        rng.standard_normal((1, yy.size)) + 3.0], axis=0)   # (6, ncell)
                    chans = np.array([0, 1, 2, 3, 4, 5 + int(cls[k])])
                    flat = (((b * A + a) * ch + chans[:, None]) * h + yy[None, :]) * h + xx[None, :]
                    per_level[li].append((flat.ravel(), vals.ravel()))
        for li, items in enumerate(per_level):
            if not items:
                continue
            flat = np.concatenate([f for f, _ in items])
            vals = np.concatenate([v for _, v in items]).astype(np.float32)
            # last writer wins, deterministically: keep the last occurrence of every index
            _, first_in_rev = np.unique(flat[::-1], return_index=True)
            keep = flat.size - 1 - first_in_rev
            idx = torch.from_numpy(flat[keep]).to(dev)
            heads[li].view(-1)[idx] = torch.from_numpy(vals[keep]).to(dev)
    return [r.view(batch, A * ch, r.shape[-2], r.shape[-1]) for r in heads]


def make_gt(batch: int, num_classes: int, size: int, lo: int, hi: int, seed: int = 0):
    """-> list (len B) of (n,6) fp32 arrays [x1,y1,x2,y2,class,mixw=1] for the training path."""
    rng = np.random.default_rng(seed + 7919)
    out = []
    for _ in range(batch):
        n = int(rng.integers(lo, hi + 1))
        boxes, cls = draw_objects(rng, n, size, num_classes)
        out.append(np.concatenate([boxes, cls[:, None].astype(np.float32),
                                   np.ones((n, 1), np.float32)], axis=1))
    return out


def make_train_heads(batch: int, num_classes: int, size: int, seed: int = 0, device="cpu",
                     strides=FPN_STRIDES, anchors_per_cell: int = 3):
    """Raw heads 0.5*N(0,1) (finite IoUs everywhere) for the loss benchmarks."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 104729)
    return [torch.randn((batch, anchors_per_cell * (5 + num_classes), size // s, size // s),
                        dtype=torch.float32, device=dev, generator=g) * 0.5 for s in strides]


def make_eval_set(n_images: int, num_classes: int, size: int = 512, seed: int = 0, gt_dtype=np.float32,
                  n_obj=(1, 12), tie_scores: bool = True):
    """Synthetic evaluation set for the AP bookkeeping (eval/evaluator.py:64-183): per image a file name, GT rows
    [x1,y1,x2,y2,class] (gt_dtype), `difficult` flags, and detections (K,6) float32 [x1,y1,x2,y2,score,class] in
    descending score like tools.torch_nms returns them: jittered copies of the GT (some duplicated, some with the
    wrong class), random false positives, and - with tie_scores - scores quantised so that exact ties occur."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_images):
        n = int(rng.integers(n_obj[0], n_obj[1] + 1))
        wh = np.exp(rng.normal(3.8, 0.8, (n, 2))).clip(6, 0.7 * size)
        c = rng.uniform(0, size, (n, 2))
        x1y1 = (c - wh / 2).clip(0, size - 3)
        x2y2 = np.maximum((c + wh / 2).clip(2, size - 1), x1y1 + 2)
        cls = rng.integers(0, num_classes, n)
        gt = np.concatenate([x1y1, x2y2, cls[:, None]], axis=1).astype(gt_dtype)
        diffs = (rng.random(n) < 0.2).astype(np.int64)
        dets = []
        for j in range(n):
            for _ in range(int(rng.integers(0, 4))):                      # 0-3 detections per object
                jit = rng.normal(0, 0.08, 4) * np.repeat(wh[j], 2)
                box = gt[j, :4].astype(np.float64) + jit
                k = cls[j] if rng.random() < 0.9 else rng.integers(0, num_classes)
                dets.append([box[0], box[1], max(box[2], box[0] + 1), max(box[3], box[1] + 1), rng.uniform(0.1, 1.0), k])
        for _ in range(int(rng.integers(0, 6))):                          # background false positives
            w, h = np.exp(rng.normal(3.5, 0.7, 2))
            x, y = rng.uniform(0, size - 10, 2)
            dets.append([x, y, x + w, y + h, rng.uniform(0.1, 0.6), rng.integers(0, num_classes)])
        d = np.array(dets, dtype=np.float32).reshape(-1, 6)
        if tie_scores and len(d):
            d[:, 4] = np.round(d[:, 4] * 20) / 20 + np.float32(0.01)      # 19 distinct scores: plenty of exact ties
        d = d[np.argsort(-d[:, 4], kind="stable")]
        out.append(("img_%05d" % i, gt, diffs, d))
    return out
