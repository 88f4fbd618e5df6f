"""TEST / BENCH INFRASTRUCTURE ONLY -- the CPU baseline, never a fallback of the product.

The reference's CPU path (predict.py:17, 33-45: Decode x3 -> cat -> recover_bboxes_prediction_* ->
per-image tools.torch_nms) restated with the same libraries the reference itself uses on a CPU box:
ATen ops for the elementwise work and torchvision.ops.batched_nms for the NMS (falling back to
oracle/nms_oracle.c where torchvision is absent).  The reference sources cannot travel to the GPU box
(/root/reference does not exist there), so this "port" is what bench.py times as cpu_baseline and as
`--impl reference`; tests/test_oracle_vs_reference.py::test_cpu_path_port_is_the_reference_sequence pins every stage
and the end result bit for bit to the live reference here (all three affines, sequential / threads / processes).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

try:
    import torchvision
    _TV = torchvision.ops.boxes.batched_nms
except Exception:          # pragma: no cover
    _TV = None


def decode_t(conv: torch.Tensor, num_classes: int, stride: int) -> torch.Tensor:
    """model/parser.py:206-235."""
    B, CH, H, W = conv.shape
    ch = 5 + num_classes
    x = conv.permute(0, 2, 3, 1).reshape(B, H, W, CH // ch, ch)
    gx = (torch.arange(W, dtype=torch.float32, device=conv.device) + 0.5).view(1, 1, W, 1, 1)
    gy = (torch.arange(H, dtype=torch.float32, device=conv.device) + 0.5).view(1, H, 1, 1, 1)
    grid = torch.cat([gx.expand(1, H, W, 1, 1), gy.expand(1, H, W, 1, 1)], dim=-1)
    # the same four views the reference splits off and one ATen call per view (model/parser.py:215-233): ATen's CPU
    # exp / sigmoid pick their vectorised or scalar inner loop from the operand's shape and strides, and the two
    # differ in the last ulp, so e.g. one sigmoid over [..., 4:] is NOT bit-identical to sigmoid(conf), sigmoid(prob)
    d1, d2, conf, prob = torch.split(x, [2, 2, 1, num_classes], dim=-1)
    xymin = (grid - torch.exp(d1)) * stride
    xymax = (grid + torch.exp(d2)) * stride
    return torch.cat((xymin, xymax, torch.sigmoid(conf), torch.sigmoid(prob)), -1)


def recover_t(pred: torch.Tensor, input_size, orig: torch.Tensor, kind: str) -> torch.Tensor:
    """dataset/base_sample.py:98-139 + the affine functions."""
    inp = torch.as_tensor(input_size, dtype=torch.float32)
    orig = orig.reshape(-1, 2)
    if kind in ("voc", "coco"):
        ratio = (inp / orig).min(dim=-1, keepdim=True)[0]
        delta = ((inp - (ratio * orig).round()) / 2).floor()
    else:
        ratio = torch.full((orig.shape[0], 1), 1.25)
        inp2 = torch.ceil(1.25 * orig / 32) * 32
        delta = ((inp2 - 1.25 * orig) / 2).floor()
    coor = (pred[..., 0:4] - delta[:, [1, 0, 1, 0]].unsqueeze(1)) / ratio.unsqueeze(1)
    edge = (orig - 1)[:, [1, 0]].unsqueeze(1)
    coor = torch.cat([coor[..., :2].clamp_min(0), torch.min(coor[..., 2:], edge)], dim=-1)
    return torch.cat([coor, pred[..., 5:] * pred[..., 4:5]], dim=-1)


def torch_nms_t(bboxes: torch.Tensor, score_threshold: float, iou_threshold: float) -> torch.Tensor:
    """tools.py:540-566."""
    scores = bboxes[:, 4:]
    mask = scores > score_threshold
    idx = mask.nonzero()
    ps, pc, pb = scores[mask], idx[:, 1], bboxes[:, :4][idx[:, 0]]
    if _TV is not None:
        keep = _TV(pb, ps, pc, iou_threshold)
    else:
        from . import pqdet_oracle as po
        keep = torch.from_numpy(po.batched_nms(pb.numpy(), ps.numpy(), pc.numpy(), iou_threshold, device="cpu"))
    if keep.numel() == 0:
        return torch.zeros((0,))
    return torch.cat([pb[keep], ps[keep, None], pc[keep, None].float()], dim=1)


def eval_chain(heads, strides, num_classes, input_size, orig, kind, thr, iou):
    """One batch through the reference's sequential CPU path.  -> list of (K,6) tensors."""
    with torch.no_grad():
        outs = [decode_t(h, num_classes, s) for h, s in zip(heads, strides)]
        pred = torch.cat([o.reshape(o.shape[0], -1, o.shape[-1]) for o in outs], dim=1)
        rec = recover_t(pred, input_size, orig, kind)
        return [torch_nms_t(rec[b], thr, iou) for b in range(rec.shape[0])]


def eval_chain_image_parallel(heads, strides, num_classes, input_size, orig, kind, thr, iou, workers: int):
    """Same work, one image per task on a thread pool with 1 intra-op thread each (torchvision's CPU NMS
    is single threaded, so this is the fair "all host cores" figure)."""
    B = heads[0].shape[0]
    orig = orig.reshape(-1, 2)

    def one(b):
        o = orig[b:b + 1] if orig.shape[0] > 1 else orig
        return eval_chain([h[b:b + 1] for h in heads], strides, num_classes, input_size, o, kind, thr, iou)[0]
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            return list(ex.map(one, range(B)))
    finally:
        torch.set_num_threads(old)


# ---- process-parallel form: the "all host cores" figure.  Threads do not give one: the per-image chain is a
# string of small ATen calls and Python glue, so a thread pool serialises on the GIL (round-1 measurement: the
# thread pool was slower than the sequential chain).  Worker processes are forked AFTER the batch exists, so
# every worker sees the heads copy-on-write (no pickling of the 1.6 MB/image inputs); each task is a contiguous
# block of images and returns its (K,6) arrays.
_PP_STATE = {}


def _pp_init():
    torch.set_num_threads(1)


def _pp_task(args):
    key, lo, hi = args
    heads, strides, num_classes, input_size, orig, kind, thr, iou = _PP_STATE[key]
    o = orig[lo:hi] if orig.shape[0] > 1 else orig
    out = eval_chain([h[lo:hi] for h in heads], strides, num_classes, input_size, o, kind, thr, iou)
    return [t.numpy() for t in out]


class ProcessPool:
    """A pool of forked workers bound to one batch; run() pushes the whole batch through the reference's CPU
    sequence, `chunk` images per task, and returns the per-image results in image order."""

    def __init__(self, heads, strides, num_classes, input_size, orig, kind, thr, iou, workers: int, chunk: int = 0):
        import multiprocessing as mp
        self.key = id(self)
        self.B = heads[0].shape[0]
        self.workers = max(1, min(workers, self.B))
        self.chunk = chunk or max(1, min(8, -(-self.B // (4 * self.workers))))
        _PP_STATE[self.key] = (heads, strides, num_classes, torch.as_tensor(input_size, dtype=torch.float32),
                               orig.reshape(-1, 2), kind, thr, iou)
        self.pool = mp.get_context("fork").Pool(self.workers, initializer=_pp_init)

    def run(self):
        tasks = [(self.key, lo, min(lo + self.chunk, self.B)) for lo in range(0, self.B, self.chunk)]
        out = []
        for part in self.pool.map(_pp_task, tasks, chunksize=1):
            out.extend(torch.from_numpy(a) for a in part)
        return out

    def close(self):
        self.pool.close()
        self.pool.join()
        _PP_STATE.pop(self.key, None)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def eval_chain_process_parallel(heads, strides, num_classes, input_size, orig, kind, thr, iou, workers: int):
    with ProcessPool(heads, strides, num_classes, input_size, orig, kind, thr, iou, workers, chunk=1) as pool:
        return pool.run()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1
