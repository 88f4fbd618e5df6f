"""Drop-ins for the IoU / NMS helpers of the reference's tools.py:335-566.

torch_nms / iou_calc3 / giou / diou / ciou take CUDA tensors and run the sm_100a kernels.
iou_calc1, iou_xywh_numpy and nms are numpy-in / numpy-out helpers of the reference
(nms has no callers there); they stage through the GPU and come back.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _ops, config


# ------------------------------------------------------------------------------------------------
# a6: tools.py:540-566
# ------------------------------------------------------------------------------------------------
_FUSED_MAX_DET = 2048            # the larger capacity class' candidate list (csrc/nms.cu FusedCfg<1>)


def _general_nms(bboxes, ids, n_sel, score_threshold, iou_threshold, nms_mode, iou_round, return_index, by_position):
    """General path until its workspace / output capacities fit.  -> det, idx, host meta (3, R) int64."""
    cap, max_det = max(n_sel * 8192, 1 << 16), 4096
    for _ in range(4):
        det, idx, meta, needed = _ops.nms_general(bboxes=bboxes, score_threshold=score_threshold,
                                                  iou_threshold=iou_threshold, nms_mode=nms_mode,
                                                  iou_round=iou_round, image_ids=ids, n_images=n_sel,
                                                  max_det=max_det, cand_capacity=cap, want_index=return_index,
                                                  out_by_position=by_position)
        host = torch.cat([meta.to(torch.int64), needed]).cpu()
        R = n_sel if by_position else bboxes.shape[0]
        counts, status = host[0:R], host[2 * R:3 * R]
        if bool((status & _lib.ST_CAND_OVERFLOW).any()):
            cap = max(int(host[3 * R]), cap * 2)
            continue
        if bool((status & _lib.ST_DET_TRUNCATED).any()):
            max_det = int(counts.max())
            continue
        return det, idx, host[:3 * R].view(3, R)
    raise _lib.PqdetError("pqdet_nms_general did not converge on a workspace size")


def batched_torch_nms(bboxes: torch.Tensor, score_threshold: float, iou_threshold: float,
                      return_index: bool = False, nms_mode: str = None, iou_round: str = None,
                      strategy: str = "auto"):
    """Batched form of torch_nms: bboxes (B, N, 4+C) -> list of B tensors (K_b, 6)
    (and, if return_index, a list of int64 tensors row*C+class).  One kernel launch and one host read for the
    whole batch (csrc/nms.cu, scores source); images whose candidates do not fit the on-chip lists are re-run
    through the general path.  strategy='general' forces the general path for every image (it is also taken for a
    negative score_threshold), 'large' picks the fused kernel's larger capacity class."""
    if bboxes.dim() != 3:
        raise ValueError("bboxes must be (B, N, 4+C)")
    B, N, _ = bboxes.shape
    C = bboxes.shape[2] - 4
    m, r = config.nms_modes()
    nms_mode, iou_round = nms_mode or m, iou_round or r
    empty = torch.zeros((0, 6), dtype=torch.float32, device=bboxes.device)
    if B == 0 or N == 0:
        e = [empty for _ in range(B)]
        return (e, [torch.zeros((0,), dtype=torch.int64, device=bboxes.device) for _ in range(B)]) \
            if return_index else e
    if strategy == "general" or score_threshold < 0:      # the fused kernel's keys order non-negative scores only
        det, idx, hm = _general_nms(bboxes, None, B, score_threshold, iou_threshold, nms_mode, iou_round,
                                    return_index, False)
        outs = [det[b, :int(hm[0, b])] for b in range(B)]
        return (outs, [idx[b, :int(hm[0, b])].to(torch.int64) for b in range(B)]) if return_index else outs
    det, idx, meta = _ops.nms_fused(bboxes, score_threshold, iou_threshold, nms_mode, iou_round, _FUSED_MAX_DET,
                                    return_index, capacity="large" if strategy == "large" else "compact")
    hm = meta[:3 * B].view(3, B).cpu()                                # the one device->host read
    outs = [det[b, :int(hm[0, b])] for b in range(B)]
    idxs = [idx[b, :int(hm[0, b])].to(torch.int64) for b in range(B)] if return_index else None
    over = torch.nonzero(hm[2] & _lib.ST_CAND_OVERFLOW).reshape(-1)
    if over.numel():
        ids = over.to(torch.int32).to(bboxes.device)
        gdet, gidx, ghm = _general_nms(bboxes, ids, int(ids.numel()), score_threshold, iou_threshold, nms_mode,
                                       iou_round, return_index, True)
        for i, b in enumerate(over.tolist()):
            k = int(ghm[0, i])
            outs[b] = gdet[i, :k]
            if return_index:
                idxs[b] = gidx[i, :k].to(torch.int64)
    return (outs, idxs) if return_index else outs


def torch_nms(bboxes: torch.Tensor, score_threshold: float, iou_threshold: float) -> torch.Tensor:
    """tools.py:540-566.  bboxes (N, 4+C) of ONE image -> (K, 6) rows [x1,y1,x2,y2,score,class] in
    descending score, or a tensor of shape (0,) when nothing is kept (tools.py:559-561)."""
    if bboxes.dim() != 2:
        raise ValueError("bboxes must be (N, 4+C)")
    out = batched_torch_nms(bboxes.unsqueeze(0), score_threshold, iou_threshold)[0]
    if out.shape[0] == 0:
        return torch.tensor([]).to(bboxes)
    return out


# ------------------------------------------------------------------------------------------------
# a7/a8: tools.py:357-477
# ------------------------------------------------------------------------------------------------
class _IouFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, b1, b2, kind):
        ctx.kind = kind
        ctx.shapes = (b1.shape, b2.shape)
        ctx.save_for_backward(b1, b2)
        return _ops.iou_pairwise(b1, b2, kind)

    @staticmethod
    def backward(ctx, g):
        b1, b2 = ctx.saved_tensors
        g1, g2 = _ops.iou_pairwise_bwd(b1, b2, g.contiguous(), ctx.kind)

        def unbroadcast(gr, shape):
            while gr.dim() > len(shape):
                gr = gr.sum(0)
            for i, s in enumerate(shape):
                if s == 1 and gr.shape[i] != 1:
                    gr = gr.sum(i, keepdim=True)
            return gr
        return unbroadcast(g1, ctx.shapes[0]), unbroadcast(g2, ctx.shapes[1]), None


def _iou(b1, b2, kind):
    if (b1.requires_grad or b2.requires_grad) and torch.is_grad_enabled():
        return _IouFn.apply(b1, b2, kind)
    return _ops.iou_pairwise(b1, b2, kind)


def iou_calc3(boxes1: torch.Tensor, boxes2: torch.Tensor):
    """tools.py:357-376 (broadcasting, last dim = x1,y1,x2,y2, no epsilon)."""
    return _iou(boxes1, boxes2, 0)


def giou(boxes1: torch.Tensor, boxes2: torch.Tensor):
    """tools.py:378-404."""
    return _iou(boxes1, boxes2, 1)


def diou(boxes1: torch.Tensor, boxes2: torch.Tensor):
    """tools.py:406-437 (adds the centre-distance term, as the reference does)."""
    return _iou(boxes1, boxes2, 2)


def ciou(boxes1: torch.Tensor, boxes2: torch.Tensor):
    """tools.py:439-477 (forward value only)."""
    return _iou(boxes1, boxes2, 3)


# ------------------------------------------------------------------------------------------------
# numpy-facing helpers
# ------------------------------------------------------------------------------------------------
def _np_to_cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def iou_calc1(boxes1: np.ndarray, boxes2: np.ndarray):
    """tools.py:335-355: numpy xyxy IoU with the union clamped at 1e-14 (fp32 on the GPU, same operation order)."""
    b1, b2 = np.asarray(boxes1), np.asarray(boxes2)
    return _ops.iou_pairwise(_np_to_cuda(b1), _np_to_cuda(b2), 4).cpu().numpy()


def iou_xywh_numpy(boxes1: np.ndarray, boxes2: np.ndarray):
    """tools.py:479-505: centre-xywh IoU (fp32 here; the mixed fp32/fp64 variant that label
    assignment needs is inside pqdet_assign_labels)."""
    def xyxy(b):
        b = np.asarray(b, dtype=np.float32)
        return np.concatenate([b[..., :2] - b[..., 2:] * 0.5, b[..., :2] + b[..., 2:] * 0.5], axis=-1)
    return _ops.iou_pairwise(_np_to_cuda(xyxy(boxes1)), _np_to_cuda(xyxy(boxes2)), 0).cpu().numpy()


def nms(bboxes, score_threshold, iou_threshold, sigma=0.3, method='nms'):
    """tools.py:507-538 (no callers in the reference).  bboxes (N,6) [x1,y1,x2,y2,score,class] -> the picked rows,
    class by class, each with the score it had when it was picked (soft-NMS decays scores in place).  Hard NMS
    ('nms') and 'soft-nms' with the reference's semantics: the first pick of a class is not compared with
    score_threshold, iou_calc1's clamped union, ties go to the earlier row.  fp32 arithmetic (the reference
    computes in the dtype of `bboxes`).  Classes come out in ascending order (the reference iterates a Python set)."""
    assert method in ['nms', 'soft-nms']
    bboxes = np.asarray(bboxes, dtype=np.float32)
    if len(bboxes) == 0:
        return np.array([])
    classes = np.unique(bboxes[:, 5])
    order = np.argsort(bboxes[:, 5], kind="stable")                   # group by class, keep the row order inside
    grouped = bboxes[order]
    seg = np.searchsorted(grouped[:, 5], classes, side="left").astype(np.int32)
    seg = np.concatenate([seg, np.array([len(grouped)], np.int32)])
    idx, score, count = _ops.classwise_nms(_np_to_cuda(grouped[:, :4]), _np_to_cuda(grouped[:, 4]),
                                           torch.from_numpy(seg).cuda(), method == 'soft-nms', float(sigma),
                                           float(score_threshold), float(iou_threshold))
    idx, score, count = idx.cpu().numpy(), score.cpu().numpy(), count.cpu().numpy()
    out = []
    for ci in range(len(classes)):
        sel = idx[seg[ci]:seg[ci] + count[ci]]
        rows = grouped[sel].copy()
        rows[:, 4] = score[seg[ci]:seg[ci] + count[ci]]
        out.append(rows)
    return np.concatenate(out, axis=0)
