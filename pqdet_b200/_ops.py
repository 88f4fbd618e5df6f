"""Thin tensor-level wrappers over the C ABI: argument checking, buffers, streams.

PyTorch is used here for device memory and streams only; all arithmetic happens in the CUDA
library.  Every function requires CUDA tensors and raises otherwise (no CPU fallback).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

_WS_CACHE = {}


def _req(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise _lib.PqdetError("%s must be a CUDA tensor: pqdet_b200 has no CPU path" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _workspace(device: torch.device, key: str, nbytes: int) -> torch.Tensor:
    """Per (device, stream, purpose) scratch, grown on demand; reuse is ordered by the stream."""
    k = (device.index, torch.cuda.current_stream(device).cuda_stream, key)
    buf = _WS_CACHE.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=device)
        _WS_CACHE[k] = buf
    return buf


def _hw_pair(x, name) -> Tuple[float, float]:
    """input_size as two host floats.  A CUDA tensor here costs a sync; pass a tuple to avoid it."""
    if isinstance(x, torch.Tensor):
        x = x.detach().reshape(-1).tolist()
    x = list(x)
    if len(x) != 2:
        raise ValueError("%s must have 2 elements (h, w)" % name)
    return float(x[0]), float(x[1])


def _orig(orig, B: int, device: torch.device) -> Tuple[torch.Tensor, int]:
    if not isinstance(orig, torch.Tensor):
        orig = torch.tensor(orig, dtype=torch.float32)
    orig = orig.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    if orig.dim() == 1 and orig.numel() == 2:
        return orig, 0
    if orig.dim() == 2 and orig.shape == (B, 2):
        return orig, 1
    raise ValueError("batch_original_size must have shape (B,2) or (2,), got %s" % (tuple(orig.shape),))


# ------------------------------------------------------------------------------------------------
def decode_fwd(raw: torch.Tensor, num_classes: int, stride: float, out: Optional[torch.Tensor] = None,
               rows_total: Optional[int] = None, row_offset: int = 0) -> torch.Tensor:
    raw = _req(raw, "conv")
    B, CH, H, W = raw.shape
    ch = 5 + num_classes
    if CH % ch:
        raise ValueError("channels (%d) not a multiple of 5+classes (%d)" % (CH, ch))
    A = CH // ch
    if out is None:
        out = torch.empty((B, H, W, A, ch), dtype=torch.float32, device=raw.device)
        rows_total, row_offset = H * W * A, 0
    _lib.check(_lib.load().pqdet_decode_fwd(_ptr(raw), _ptr(out), B, A, num_classes, H, W, float(stride),
                                            int(rows_total), int(row_offset), _dev(raw), _stream(raw.device)),
               "pqdet_decode_fwd")
    return out


def decode_levels(raws: Sequence[torch.Tensor], num_classes: int, strides: Sequence[float]) -> torch.Tensor:
    """Every level decoded straight into its row range of (B, N, 5+C): one launch (pqdet_decode_levels)."""
    raws = [_req(r, "conv") for r in raws]
    ch = 5 + num_classes
    B = raws[0].shape[0]
    A = raws[0].shape[1] // ch
    for r in raws:
        if r.shape[0] != B or r.shape[1] != A * ch or r.device != raws[0].device:
            raise ValueError("heads have inconsistent shapes/devices")
    L = len(raws)
    N = sum(r.shape[2] * r.shape[3] * A for r in raws)
    out = torch.empty((B, N, ch), dtype=torch.float32, device=raws[0].device)
    VP, IP, FP = ctypes.c_void_p * L, ctypes.c_int * L, ctypes.c_float * L
    _lib.check(_lib.load().pqdet_decode_levels(L, VP(*[r.data_ptr() for r in raws]), IP(*[r.shape[2] for r in raws]),
                                               IP(*[r.shape[3] for r in raws]), FP(*[float(s) for s in strides]),
                                               _ptr(out), B, A, num_classes, _dev(raws[0]), _stream(raws[0].device)),
               "pqdet_decode_levels")
    return out


def head_conv_decode(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], num_classes: int,
                     stride: float, out: Optional[torch.Tensor] = None, rows_total: Optional[int] = None,
                     row_offset: int = 0, want_raw: bool = False, want_decoded: bool = True):
    """1x1 head convolution + Decode on the tensor cores (pqdet_head_conv_decode).
    x (B,Cin,H,W), weight (A*(5+C), Cin[,1,1]), bias (A*(5+C))|None -> decoded (B,H,W,A,5+C) [or rows of `out`],
    raw (B, A*(5+C), H, W) if want_raw; want_decoded=False (with want_raw): only the raw head."""
    x = _req(x, "x")
    weight = _req(weight.reshape(weight.shape[0], -1), "weight")
    if bias is not None:
        bias = _req(bias, "bias")
    B, Cin, H, W = x.shape
    ch = 5 + num_classes
    ACH = weight.shape[0]
    if ACH % ch or weight.shape[1] != Cin:
        raise ValueError("weight %s does not match Cin=%d / 5+classes=%d" % (tuple(weight.shape), Cin, ch))
    A = ACH // ch
    if not want_decoded:
        if not want_raw:
            raise ValueError("nothing to compute")
        out, rows_total, row_offset = None, H * W * A, 0
    elif out is None:
        out = torch.empty((B, H, W, A, ch), dtype=torch.float32, device=x.device)
        rows_total, row_offset = H * W * A, 0
    raw = torch.empty((B, ACH, H, W), dtype=torch.float32, device=x.device) if want_raw else None
    _lib.check(_lib.load().pqdet_head_conv_decode(_ptr(x), _ptr(weight), _ptr(bias), _ptr(out), _ptr(raw), B, Cin, H, W,
                                                  A, num_classes, float(stride), int(rows_total), int(row_offset),
                                                  _dev(x), _stream(x.device)), "pqdet_head_conv_decode")
    if not want_decoded:
        return raw
    return (out, raw) if want_raw else out


def head_conv_decode_levels(xs: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                            biases: Sequence[Optional[torch.Tensor]], num_classes: int, strides: Sequence[float],
                            out: torch.Tensor) -> bool:
    """All levels' head convolution + Decode in one launch into `out` (B, sum rows, 5+C)
    (pqdet_head_conv_decode_levels).  Returns False when a level does not qualify for the persistent kernel (the
    caller then runs head_conv_decode per level)."""
    xs = [_req(x, "x") for x in xs]
    ws = [_req(w.reshape(w.shape[0], -1), "weight") for w in weights]
    bs = [None if b is None else _req(b, "bias") for b in biases]
    L = len(xs)
    ch = 5 + num_classes
    B = xs[0].shape[0]
    A = ws[0].shape[0] // ch
    for x, w in zip(xs, ws):
        if x.shape[0] != B or w.shape[0] != A * ch or w.shape[1] != x.shape[1] or x.device != xs[0].device:
            raise ValueError("features / weights have inconsistent shapes or devices")
    if out.shape != (B, sum(x.shape[2] * x.shape[3] * A for x in xs), ch) or not out.is_contiguous():
        raise ValueError("out must be a contiguous (B, sum_l H_l*W_l*A, 5+C) tensor")
    VP, IP, FP = ctypes.c_void_p * L, ctypes.c_int * L, ctypes.c_float * L
    rc = _lib.load().pqdet_head_conv_decode_levels(
        L, VP(*[x.data_ptr() for x in xs]), VP(*[w.data_ptr() for w in ws]),
        VP(*[None if b is None else b.data_ptr() for b in bs]), IP(*[x.shape[1] for x in xs]),
        IP(*[x.shape[2] for x in xs]), IP(*[x.shape[3] for x in xs]), FP(*[float(s) for s in strides]), _ptr(out), B, A,
        num_classes, _dev(xs[0]), _stream(xs[0].device))
    if rc == _lib.PQDET_ERR_UNSUPPORTED:
        return False
    _lib.check(rc, "pqdet_head_conv_decode_levels")
    return True


def decode_bwd(raw: torch.Tensor, grad_out: torch.Tensor, num_classes: int, stride: float) -> torch.Tensor:
    raw = _req(raw, "conv")
    grad_out = _req(grad_out, "grad_out")
    B, CH, H, W = raw.shape
    A = CH // (5 + num_classes)
    grad = torch.empty_like(raw)
    _lib.check(_lib.load().pqdet_decode_bwd(_ptr(raw), _ptr(grad_out), _ptr(grad), B, A, num_classes, H, W,
                                            float(stride), H * W * A, 0, _dev(raw), _stream(raw.device)),
               "pqdet_decode_bwd")
    return grad


def recover(pred: torch.Tensor, input_size, original_size, kind: str) -> torch.Tensor:
    pred = _req(pred, "batch_pred_bbox")
    if pred.dim() != 3:
        raise ValueError("batch_pred_bbox must be (B, N, 5+C)")
    B, N, ch = pred.shape
    C = ch - 5
    in_h, in_w = _hw_pair(input_size, "input_size")
    orig, per = _orig(original_size, B, pred.device)
    out = torch.empty((B, N, 4 + C), dtype=torch.float32, device=pred.device)
    _lib.check(_lib.load().pqdet_recover(_ptr(pred), _ptr(out), B, N, C, _lib.AFFINE[kind], in_h, in_w,
                                         _ptr(orig), per, _dev(pred), _stream(pred.device)), "pqdet_recover")
    return out


def make_heads(raws: Sequence[torch.Tensor], strides: Sequence[float], num_classes: int, input_size,
               original_size, kind: str, score_threshold: float, iou_threshold: float,
               nms_mode: str, iou_round: str):
    raws = [_req(r, "head") for r in raws]
    if not 1 <= len(raws) <= _lib.MAX_LEVELS or len(raws) != len(strides):
        raise ValueError("need 1..%d heads with matching strides" % _lib.MAX_LEVELS)
    B = raws[0].shape[0]
    ch = 5 + num_classes
    A = raws[0].shape[1] // ch
    h = _lib.HeadsT()
    for i, (r, s) in enumerate(zip(raws, strides)):
        if r.shape[0] != B or r.shape[1] != A * ch or r.device != raws[0].device:
            raise ValueError("head %d has inconsistent shape/device" % i)
        h.raw[i] = r.data_ptr()
        h.H[i], h.W[i] = int(r.shape[2]), int(r.shape[3])
        h.stride[i] = float(s)
    h.n_levels = len(raws)
    h.B, h.A, h.C = B, A, num_classes
    h.affine_kind = _lib.AFFINE[kind]
    h.in_h, h.in_w = _hw_pair(input_size, "input_size")
    orig, per = _orig(original_size, B, raws[0].device)
    h.orig_hw = orig.data_ptr()
    h.orig_per_image = per
    h.score_threshold = float(score_threshold)
    h.iou_threshold = float(iou_threshold)
    h.nms_mode = _lib.NMS_MODE[nms_mode]
    h.iou_round = _lib.IOU_ROUND[iou_round]
    keep_alive = (raws, orig)
    return h, keep_alive


def alloc_fused_outputs(B: int, max_det: int, want_index: bool, device):
    """Output buffers of the fused kernel.  meta = [counts(B), ncand(B), status(B), scheduler words(2), images done(1),
    pad(1)]."""
    det = torch.empty((B, max_det, 6), dtype=torch.float32, device=device)
    idx = torch.empty((B, max_det), dtype=torch.int32, device=device) if want_index else None
    meta = torch.zeros((3 * B + 4,), dtype=torch.int32, device=device)
    return det, idx, meta


_FUSED_ARMED = set()


def decode_nms_fused(heads_t, keep_alive, max_det: int, want_index: bool, out=None, capacity: str = "compact"):
    """-> det (B,max_det,6), idx (B,max_det)|None, meta int32 (3B+2): counts, ncand, status, scheduler words.
    `out` = buffers from alloc_fused_outputs to reuse across calls: no allocation in the hot loop, and from
    the second call on nothing but the kernel is enqueued (the kernel re-arms its own scheduler words).
    capacity: 'compact' (512 hit rows / 1280 candidates per image on chip, 7 CTAs per SM) or 'large' (1024 / 2048)."""
    raws, _ = keep_alive
    device = raws[0].device
    B = heads_t.B
    det, idx, meta = out if out is not None else alloc_fused_outputs(B, max_det, want_index, device)
    counts, ncand, status, work = meta[0:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:]
    key = (work.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    armed = 1 if (out is not None and key in _FUSED_ARMED) else 0
    _FUSED_ARMED.discard(key)                       # if the call raises, the next one zeroes the words again
    _lib.check(_lib.load().pqdet_decode_nms(ctypes.byref(heads_t), _ptr(det), _ptr(idx), int(max_det),
                                            _ptr(counts), _ptr(ncand), _ptr(status), _ptr(work), armed,
                                            _lib.CAPACITY[capacity], _dev(raws[0]), _stream(device)),
               "pqdet_decode_nms")
    if out is not None:
        _FUSED_ARMED.add(key)
    return det, idx, meta


def _req_pinned(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if t.is_cuda:
        return _req(t, name)
    if not t.is_pinned():
        raise _lib.PqdetError("%s is pageable host memory: the host-buffer path needs page-locked tensors "
                              "(tensor.pin_memory()); pqdet_b200 has no CPU path" % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("%s must be contiguous float32" % name)
    return t


def make_heads_host(raws: Sequence[torch.Tensor], strides: Sequence[float], num_classes: int, input_size,
                    original_size, kind: str, score_threshold: float, iou_threshold: float,
                    nms_mode: str, iou_round: str):
    """make_heads for head tensors that live in PINNED HOST memory (no staging copy is made)."""
    raws = [_req_pinned(r, "head") for r in raws]
    if not 1 <= len(raws) <= _lib.MAX_LEVELS or len(raws) != len(strides):
        raise ValueError("need 1..%d heads with matching strides" % _lib.MAX_LEVELS)
    B = raws[0].shape[0]
    ch = 5 + num_classes
    A = raws[0].shape[1] // ch
    h = _lib.HeadsT()
    for i, (r, s) in enumerate(zip(raws, strides)):
        if r.shape[0] != B or r.shape[1] != A * ch:
            raise ValueError("head %d has inconsistent shape" % i)
        h.raw[i] = r.data_ptr()
        h.H[i], h.W[i] = int(r.shape[2]), int(r.shape[3])
        h.stride[i] = float(s)
    h.n_levels = len(raws)
    h.B, h.A, h.C = B, A, num_classes
    h.affine_kind = _lib.AFFINE[kind]
    h.in_h, h.in_w = _hw_pair(input_size, "input_size")
    if not isinstance(original_size, torch.Tensor):
        original_size = torch.tensor(original_size, dtype=torch.float32)
    orig = original_size.detach().to(dtype=torch.float32).contiguous()
    if not orig.is_cuda and not orig.is_pinned():
        orig = orig.pin_memory()
    if orig.dim() == 1 and orig.numel() == 2:
        per = 0
    elif orig.dim() == 2 and tuple(orig.shape) == (B, 2):
        per = 1
    else:
        raise ValueError("batch_original_size must have shape (B,2) or (2,), got %s" % (tuple(orig.shape),))
    h.orig_hw = orig.data_ptr()
    h.orig_per_image = per
    h.score_threshold = float(score_threshold)
    h.iou_threshold = float(iou_threshold)
    h.nms_mode = _lib.NMS_MODE[nms_mode]
    h.iou_round = _lib.IOU_ROUND[iou_round]
    return h, (raws, orig)


def alloc_host_outputs(B: int, max_det: int, want_index: bool, device):
    """Pinned-host outputs of decode_nms_host (+ the two device scheduler words)."""
    det = torch.empty((B, max_det, 6), dtype=torch.float32, pin_memory=True)
    idx = torch.empty((B, max_det), dtype=torch.int32, pin_memory=True) if want_index else None
    meta = torch.zeros((3 * B,), dtype=torch.int32, pin_memory=True)
    work = torch.zeros((2,), dtype=torch.int32, device=device)
    return det, idx, meta, work


def decode_nms_host(heads_t, keep_alive, max_det: int, want_index: bool, device, out=None,
                    capacity: str = "large"):
    """Host buffers in, host buffers out (pqdet_decode_nms_host).  -> det, idx, meta (pinned host), work (device).
    Asynchronous on the current stream of `device`: synchronise before reading the outputs."""
    device = torch.device(device)
    B = heads_t.B
    det, idx, meta, work = out if out is not None else alloc_host_outputs(B, max_det, want_index, device)
    counts, ncand, status = meta[0:B], meta[B:2 * B], meta[2 * B:3 * B]
    key = (work.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    armed = 1 if (out is not None and key in _FUSED_ARMED) else 0
    _FUSED_ARMED.discard(key)
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    _lib.check(_lib.load().pqdet_decode_nms_host(ctypes.byref(heads_t), _ptr(det), _ptr(idx), int(max_det),
                                                 _ptr(counts), _ptr(ncand), _ptr(status), _ptr(work), armed,
                                                 _lib.CAPACITY[capacity], dev_index, _stream(device)),
               "pqdet_decode_nms_host")
    if out is not None:
        _FUSED_ARMED.add(key)
    return det, idx, meta, work


def nms_fused(bboxes: torch.Tensor, score_threshold: float, iou_threshold: float, nms_mode: str, iou_round: str,
              max_det: int, want_index: bool, out=None, capacity: str = "compact"):
    """tools.torch_nms for a batch in one launch.  bboxes (B, N, 4+C) -> det, idx, meta (as decode_nms_fused)."""
    bboxes = _req(bboxes, "bboxes")
    if bboxes.dim() != 3:
        raise ValueError("bboxes must be (B, N, 4+C)")
    B, N, C = bboxes.shape[0], bboxes.shape[1], bboxes.shape[2] - 4
    device = bboxes.device
    det, idx, meta = out if out is not None else alloc_fused_outputs(B, max_det, want_index, device)
    counts, ncand, status, work = meta[0:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:]
    key = (work.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    armed = 1 if (out is not None and key in _FUSED_ARMED) else 0
    _FUSED_ARMED.discard(key)
    _lib.check(_lib.load().pqdet_nms_fused(_ptr(bboxes), B, N, C, float(score_threshold), float(iou_threshold),
                                           _lib.NMS_MODE[nms_mode], _lib.IOU_ROUND[iou_round], _ptr(det),
                                           _ptr(idx), int(max_det), _ptr(counts), _ptr(ncand), _ptr(status),
                                           _ptr(work), armed, _lib.CAPACITY[capacity], _dev(bboxes), _stream(device)),
               "pqdet_nms_fused")
    if out is not None:
        _FUSED_ARMED.add(key)
    return det, idx, meta


def nms_general(heads_t=None, keep_alive=None, bboxes: Optional[torch.Tensor] = None,
                score_threshold: float = 0.0, iou_threshold: float = 0.0, nms_mode: str = "auto_cuda",
                iou_round: str = "tv_cuda", image_ids: Optional[torch.Tensor] = None,
                n_images: Optional[int] = None, max_det: int = 4096, cand_capacity: int = 1 << 16,
                want_index: bool = False, out_by_position: bool = False):
    """One attempt of the general path.  Outputs are indexed by image id, or by position in
    image_ids (and sized n_images) when out_by_position is set.
    -> det, idx, meta (3,B) int32 [counts, ncand, status], needed (1,) int64 (device)."""
    lib = _lib.load()
    if heads_t is not None:
        device = keep_alive[0][0].device
        B, C, N = heads_t.B, heads_t.C, 0
        for i in range(heads_t.n_levels):
            N += heads_t.H[i] * heads_t.W[i] * heads_t.A
        bptr, hptr, from_heads = None, ctypes.byref(heads_t), 1
    else:
        bboxes = _req(bboxes, "bboxes")
        if bboxes.dim() != 3:
            raise ValueError("bboxes must be (B, N, 4+C)")
        device = bboxes.device
        B, N, C = bboxes.shape[0], bboxes.shape[1], bboxes.shape[2] - 4
        bptr, hptr, from_heads = _ptr(bboxes), None, 0
    if n_images is None:
        n_images = B if image_ids is None else int(image_ids.numel())
    wbytes = lib.pqdet_nms_general_workspace(n_images, N, C, int(cand_capacity), from_heads)
    if wbytes < 0:
        _lib.check(int(wbytes), "pqdet_nms_general_workspace")
    ws = _workspace(device, "nms_general", wbytes)
    R = n_images if out_by_position else B
    det = torch.empty((R, max_det, 6), dtype=torch.float32, device=device)
    idx = torch.empty((R, max_det), dtype=torch.int32, device=device) if want_index else None
    meta = torch.zeros((3 * R,), dtype=torch.int32, device=device)
    needed = torch.zeros((1,), dtype=torch.int64, device=device)
    counts, ncand, status = meta[0:R], meta[R:2 * R], meta[2 * R:3 * R]
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    _lib.check(lib.pqdet_nms_general(hptr, bptr, N, B, C, float(score_threshold), float(iou_threshold),
                                     _lib.NMS_MODE[nms_mode], _lib.IOU_ROUND[iou_round],
                                     _ptr(image_ids), int(n_images), 1 if out_by_position else 0,
                                     _ptr(det), _ptr(idx), int(max_det),
                                     _ptr(counts), _ptr(ncand), _ptr(status), _ptr(ws), int(ws.numel()),
                                     int(cand_capacity), _ptr(needed), dev_index, _stream(device)),
               "pqdet_nms_general")
    return det, idx, meta, needed


def iou_pairwise(b1: torch.Tensor, b2: torch.Tensor, kind: int) -> torch.Tensor:
    b1, b2 = torch.broadcast_tensors(b1, b2)
    b1, b2 = _req(b1, "boxes1"), _req(b2, "boxes2")
    if b1.shape[-1] != 4:
        raise ValueError("last dimension must be 4 (x1,y1,x2,y2)")
    out = torch.empty(b1.shape[:-1], dtype=torch.float32, device=b1.device)
    n = out.numel()
    _lib.check(_lib.load().pqdet_iou_pairwise(_ptr(b1), _ptr(b2), _ptr(out), n, kind, _dev(b1),
                                              _stream(b1.device)), "pqdet_iou_pairwise")
    return out


def iou_pairwise_bwd(b1, b2, grad_out, kind: int):
    b1, b2 = torch.broadcast_tensors(b1, b2)
    b1, b2 = _req(b1, "boxes1"), _req(b2, "boxes2")
    grad_out = _req(grad_out, "grad_out")
    g1, g2 = torch.empty_like(b1), torch.empty_like(b2)
    _lib.check(_lib.load().pqdet_iou_pairwise_bwd(_ptr(b1), _ptr(b2), _ptr(grad_out), _ptr(g1), _ptr(g2),
                                                  grad_out.numel(), kind, _dev(b1), _stream(b1.device)),
               "pqdet_iou_pairwise_bwd")
    return g1, g2


def loss_fwd_bwd(x: torch.Tensor, input_is_raw: bool, label: torch.Tensor, gt: torch.Tensor,
                 num_classes: int, stride: float, bbox_loss: str, ignore_thresh: float,
                 l1_loss_gain: float, want_grad: bool):
    """-> out4 (4,) device [loss,bbox,conf,cls], nan_flag (1,) int32 device, grad|None."""
    x, label, gt = _req(x, "pred"), _req(label, "label"), _req(gt, "bboxes")
    C = num_classes
    if bbox_loss == "ciou":
        # tools.py:472 / model/loss.py:110-114: atan(0/0) at empty label cells makes the reference
        # raise on every call; parity for ciou is "raises".
        raise RuntimeError("NaN in loss")
    if bbox_loss not in _lib.BBOX_LOSS:
        raise NotImplementedError(bbox_loss)
    if input_is_raw:
        B, CH, H, W = x.shape
        A = CH // (5 + C)
    else:
        B, H, W, A, _ = x.shape
    if tuple(label.shape) != (B, H, W, A, 6 + C):
        raise ValueError("label shape %s != %s" % (tuple(label.shape), (B, H, W, A, 6 + C)))
    if gt.dim() != 3 or gt.shape[0] != B or gt.shape[2] != 4 or gt.shape[1] < 1:
        raise ValueError("bboxes must be (B, G>=1, 4)")
    lib = _lib.load()
    ws = _workspace(x.device, "loss", lib.pqdet_loss_workspace(B, A, H, W))
    grad = torch.empty_like(x) if want_grad else None
    out = torch.empty((4,), dtype=torch.float32, device=x.device)
    flag = torch.empty((1,), dtype=torch.int32, device=x.device)
    _lib.check(lib.pqdet_loss_fwd_bwd(_ptr(x), 1 if input_is_raw else 0, _ptr(label), _ptr(gt), _ptr(grad),
                                      _ptr(out), _ptr(flag), _ptr(ws), B, A, C, H, W, int(gt.shape[1]),
                                      float(stride), _lib.BBOX_LOSS[bbox_loss], float(ignore_thresh),
                                      float(l1_loss_gain), _dev(x), _stream(x.device)), "pqdet_loss_fwd_bwd")
    return out, flag, grad


def loss_scale_grad(grad: torch.Tensor, input_is_raw: bool, num_classes: int, g_loss, g_bbox, g_conf, g_cls):
    """In-place per-channel-group chain rule (see pqdet_loss_scale_grad)."""
    if input_is_raw:
        B, CH, H, W = grad.shape
        A = CH // (5 + num_classes)
    else:
        B, H, W, A, _ = grad.shape
    gs = [None if g is None else _req(g.reshape(-1)[:1], "upstream grad") for g in (g_loss, g_bbox, g_conf, g_cls)]
    _lib.check(_lib.load().pqdet_loss_scale_grad(_ptr(grad), 1 if input_is_raw else 0, B, A, num_classes, H, W,
                                                 _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]), _ptr(gs[3]),
                                                 _dev(grad), _stream(grad.device)), "pqdet_loss_scale_grad")
    return grad




def loss_levels(raws, labels, gts, num_classes: int, strides, bbox_loss: str, ignore_thresh: float,
                l1_loss_gain: float, want_grad: bool):
    """All levels in one launch.  -> out (4+5L,) device, nan_flag (1,) int32, grads list|None."""
    if bbox_loss == "ciou":
        raise RuntimeError("NaN in loss")          # the reference always raises for ciou
    if bbox_loss not in _lib.BBOX_LOSS:
        raise NotImplementedError(bbox_loss)
    raws = [_req(r, "head") for r in raws]
    labels = [_req(x, "label") for x in labels]
    gts = [_req(x, "bboxes") for x in gts]
    L = len(raws)
    C = num_classes
    B = raws[0].shape[0]
    A = raws[0].shape[1] // (5 + C)
    device = raws[0].device
    for r, lab, g in zip(raws, labels, gts):
        _, CH, H, W = r.shape
        if r.shape[0] != B or CH != A * (5 + C) or tuple(lab.shape) != (B, H, W, A, 6 + C):
            raise ValueError("head/label shapes do not match: %s vs %s" % (tuple(r.shape), tuple(lab.shape)))
        if g.dim() != 3 or g.shape[0] != B or g.shape[2] != 4 or g.shape[1] < 1:
            raise ValueError("bboxes must be (B, G>=1, 4)")
    grads = [torch.empty_like(r) for r in raws] if want_grad else None
    VP = ctypes.c_void_p * L
    IP = ctypes.c_int * L
    Hs, Ws = IP(*[r.shape[2] for r in raws]), IP(*[r.shape[3] for r in raws])
    lib = _lib.load()
    ws = _workspace(device, "loss_levels", lib.pqdet_loss_levels_workspace(L, B, A, Hs, Ws))
    out = torch.empty((4 + 5 * L,), dtype=torch.float32, device=device)
    flag = torch.empty((1,), dtype=torch.int32, device=device)
    _lib.check(lib.pqdet_loss_levels(
        L, VP(*[r.data_ptr() for r in raws]), VP(*[x.data_ptr() for x in labels]), VP(*[x.data_ptr() for x in gts]),
        VP(*[g.data_ptr() for g in grads]) if want_grad else None, Hs, Ws, IP(*[g.shape[1] for g in gts]),
        (ctypes.c_float * L)(*[float(s) for s in strides]), B, A, C, _lib.BBOX_LOSS[bbox_loss],
        float(ignore_thresh), float(l1_loss_gain), _ptr(out), _ptr(flag), _ptr(ws),
        1, _dev(raws[0]), _stream(device)), "pqdet_loss_levels")
    return out, flag, grads


def loss_levels_scale_grad(grads, num_classes: int, upstream: torch.Tensor):
    L = len(grads)
    B, CH = grads[0].shape[0], grads[0].shape[1]
    A = CH // (5 + num_classes)
    upstream = _req(upstream, "upstream grad")
    VP = ctypes.c_void_p * L
    IP = ctypes.c_int * L
    _lib.check(_lib.load().pqdet_loss_levels_scale_grad(
        L, VP(*[g.data_ptr() for g in grads]), IP(*[g.shape[2] for g in grads]), IP(*[g.shape[3] for g in grads]),
        B, A, num_classes, _ptr(upstream), _dev(grads[0]), _stream(grads[0].device)), "pqdet_loss_levels_scale_grad")
    return grads


def assign_labels(gt: torch.Tensor, gt_count: torch.Tensor, num_classes: int, anchors, strides, sizes_hw,
                  iou_threshold: float, list_capacity: Optional[int] = None):
    """gt (B,n_max,6) cuda, gt_count (B) int32 cuda -> labels[3], gtlists[3] (B,cap,4), list_len (B,3)."""
    gt = _req(gt, "gt")
    if gt.dim() != 3 or gt.shape[2] != 6:
        raise ValueError("gt must be (B, n_max, 6)")
    B, n_max = gt.shape[0], gt.shape[1]
    device = gt.device
    gt_count = gt_count.to(device=device, dtype=torch.int32).contiguous()
    C = num_classes
    anc = (ctypes.c_float * 18)(*[float(v) for wh in anchors for v in wh])
    st = (ctypes.c_int * 3)(*[int(s) for s in strides])
    Hs = (ctypes.c_int * 3)(*[int(s[0]) for s in sizes_hw])
    Ws = (ctypes.c_int * 3)(*[int(s[1]) for s in sizes_hw])
    cap = int(list_capacity) if list_capacity else max(3 * n_max, 1)
    # one allocation for the three label tensors (and one for the three lists): the library then fills the
    # background with a single streaming launch
    sizes = [B * Hs[i] * Ws[i] * 3 * (6 + C) for i in range(3)]
    lbuf = torch.empty((sum(sizes),), dtype=torch.float32, device=device)
    labels, off = [], 0
    for i in range(3):
        labels.append(lbuf[off:off + sizes[i]].view(B, Hs[i], Ws[i], 3, 6 + C))
        off += sizes[i]
    lists = list(torch.empty((3, B, cap, 4), dtype=torch.float32, device=device).unbind(0))
    list_len = torch.empty((B, 3), dtype=torch.int32, device=device)
    lib = _lib.load()
    owner = _workspace(device, "assign", lib.pqdet_assign_workspace(B, Hs, Ws))
    _lib.check(lib.pqdet_assign_labels(_ptr(gt), _ptr(gt_count), B, n_max, C, anc, st, Hs, Ws,
                                       float(iou_threshold), _ptr(labels[0]), _ptr(labels[1]), _ptr(labels[2]),
                                       _ptr(lists[0]), _ptr(lists[1]), _ptr(lists[2]), cap, _ptr(list_len),
                                       _ptr(owner), _dev(gt), _stream(device)), "pqdet_assign_labels")
    return labels, lists, list_len


def assign_sparse(gt: torch.Tensor, gt_count: torch.Tensor, anchors, strides, sizes_hw, iou_threshold: float,
                  list_capacity: Optional[int] = None):
    """gt (B,n_max,6) cuda, gt_count (B) -> owner maps [3] int32 (B,3,H_s,W_s), gtlists [3] (B,cap,4), list_len (B,3).
    The dense label tensors are never produced (pqdet_assign_sparse)."""
    gt = _req(gt, "gt")
    if gt.dim() != 3 or gt.shape[2] != 6:
        raise ValueError("gt must be (B, n_max, 6)")
    B, n_max = gt.shape[0], gt.shape[1]
    device = gt.device
    gt_count = gt_count.to(device=device, dtype=torch.int32).contiguous()
    anc = (ctypes.c_float * 18)(*[float(v) for wh in anchors for v in wh])
    st = (ctypes.c_int * 3)(*[int(s) for s in strides])
    Hs = (ctypes.c_int * 3)(*[int(s[0]) for s in sizes_hw])
    Ws = (ctypes.c_int * 3)(*[int(s[1]) for s in sizes_hw])
    cap = int(list_capacity) if list_capacity else max(3 * n_max, 1)
    sizes = [B * 3 * Hs[i] * Ws[i] for i in range(3)]
    obuf = torch.empty((sum(sizes),), dtype=torch.int32, device=device)
    owners, off = [], 0
    for i in range(3):
        owners.append(obuf[off:off + sizes[i]].view(B, 3, Hs[i], Ws[i]))
        off += sizes[i]
    lists = list(torch.empty((3, B, cap, 4), dtype=torch.float32, device=device).unbind(0))
    list_len = torch.empty((B, 3), dtype=torch.int32, device=device)
    _lib.check(_lib.load().pqdet_assign_sparse(_ptr(gt), _ptr(gt_count), B, n_max, anc, st, Hs, Ws,
                                               float(iou_threshold), _ptr(owners[0]), _ptr(owners[1]),
                                               _ptr(owners[2]), _ptr(lists[0]), _ptr(lists[1]), _ptr(lists[2]), cap,
                                               _ptr(list_len), _dev(gt), _stream(device)), "pqdet_assign_sparse")
    return owners, lists, list_len


def loss_levels_sparse(raws, owners, gt6: torch.Tensor, gts, num_classes: int, strides, bbox_loss: str,
                       ignore_thresh: float, l1_loss_gain: float, want_grad: bool):
    """loss_levels with sparse targets: owners[l] (B,3,H_l,W_l) int32, gt6 (B,n_max,6), gts[l] (B,G_l,4)."""
    if bbox_loss == "ciou":
        raise RuntimeError("NaN in loss")
    if bbox_loss not in _lib.BBOX_LOSS:
        raise NotImplementedError(bbox_loss)
    raws = [_req(r, "head") for r in raws]
    gts = [_req(x, "bboxes") for x in gts]
    gt6 = _req(gt6, "gt")
    L = len(raws)
    C = num_classes
    B = raws[0].shape[0]
    A = raws[0].shape[1] // (5 + C)
    device = raws[0].device
    for r, o, g in zip(raws, owners, gts):
        _, CH, H, W = r.shape
        if r.shape[0] != B or CH != A * (5 + C) or tuple(o.shape) != (B, A, H, W) or o.dtype != torch.int32 \
                or not o.is_cuda or not o.is_contiguous():
            raise ValueError("head/owner shapes do not match: %s vs %s" % (tuple(r.shape), tuple(o.shape)))
        if g.dim() != 3 or g.shape[0] != B or g.shape[2] != 4 or g.shape[1] < 1:
            raise ValueError("bboxes must be (B, G>=1, 4)")
    if gt6.dim() != 3 or gt6.shape[0] != B or gt6.shape[2] != 6:
        raise ValueError("gt must be (B, n_max, 6)")
    grads = [torch.empty_like(r) for r in raws] if want_grad else None
    VP = ctypes.c_void_p * L
    IP = ctypes.c_int * L
    Hs, Ws = IP(*[r.shape[2] for r in raws]), IP(*[r.shape[3] for r in raws])
    lib = _lib.load()
    ws = _workspace(device, "loss_levels", lib.pqdet_loss_levels_workspace(L, B, A, Hs, Ws))
    out = torch.empty((4 + 5 * L,), dtype=torch.float32, device=device)
    flag = torch.empty((1,), dtype=torch.int32, device=device)
    _lib.check(lib.pqdet_loss_levels_sparse(
        L, VP(*[r.data_ptr() for r in raws]), VP(*[o.data_ptr() for o in owners]), _ptr(gt6), int(gt6.shape[1]),
        VP(*[x.data_ptr() for x in gts]), VP(*[g.data_ptr() for g in grads]) if want_grad else None, Hs, Ws,
        IP(*[g.shape[1] for g in gts]), (ctypes.c_float * L)(*[float(s) for s in strides]), B, A, C,
        _lib.BBOX_LOSS[bbox_loss], float(ignore_thresh), float(l1_loss_gain), _ptr(out), _ptr(flag), _ptr(ws),
        1, _dev(raws[0]), _stream(device)), "pqdet_loss_levels_sparse")
    return out, flag, grads


def classwise_nms(boxes: torch.Tensor, scores: torch.Tensor, seg_off: torch.Tensor, soft: bool, sigma: float,
                  score_threshold: float, iou_threshold: float):
    """tools.nms on class-grouped rows (pqdet_classwise_nms).  -> picked indices (n), their scores (n), counts (C)."""
    boxes, scores = _req(boxes, "boxes"), _req(scores, "scores").clone()
    n, n_classes = boxes.shape[0], seg_off.numel() - 1
    dev = boxes.device
    seg_off = seg_off.to(device=dev, dtype=torch.int32).contiguous()
    out_idx = torch.zeros((max(n, 1),), dtype=torch.int32, device=dev)
    out_score = torch.zeros((max(n, 1),), dtype=torch.float32, device=dev)
    out_count = torch.zeros((max(n_classes, 1),), dtype=torch.int32, device=dev)
    alive = torch.empty((max(n, 1),), dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().pqdet_classwise_nms(_ptr(boxes), _ptr(scores), _ptr(seg_off), int(n_classes), int(n),
                                               1 if soft else 0, float(sigma), float(score_threshold),
                                               float(iou_threshold), _ptr(out_idx), _ptr(out_score), _ptr(out_count),
                                               _ptr(alive), _dev(boxes), _stream(dev)), "pqdet_classwise_nms")
    return out_idx[:n], out_score[:n], out_count[:n_classes]


def make_geometry(B: int, A: int, shapes: Sequence[Tuple[int, int]], strides: Sequence[float], num_classes: int,
                  input_size, original_size, kind: str, score_threshold: float, iou_threshold: float,
                  nms_mode: str, iou_round: str, device):
    """pqdet_heads_t without raw tensors (pqdet_records_nms only needs the level geometry)."""
    if not 1 <= len(shapes) <= _lib.MAX_LEVELS or len(shapes) != len(strides):
        raise ValueError("need 1..%d levels with matching strides" % _lib.MAX_LEVELS)
    h = _lib.HeadsT()
    for i, ((H, W), s) in enumerate(zip(shapes, strides)):
        h.raw[i] = None
        h.H[i], h.W[i] = int(H), int(W)
        h.stride[i] = float(s)
    h.n_levels = len(shapes)
    h.B, h.A, h.C = B, A, num_classes
    h.affine_kind = _lib.AFFINE[kind]
    h.in_h, h.in_w = _hw_pair(input_size, "input_size")
    orig, per = _orig(original_size, B, device)
    h.orig_hw = orig.data_ptr()
    h.orig_per_image = per
    h.score_threshold = float(score_threshold)
    h.iou_threshold = float(iou_threshold)
    h.nms_mode = _lib.NMS_MODE[nms_mode]
    h.iou_round = _lib.IOU_ROUND[iou_round]
    return h, (orig,)


def head_conv_hits(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], num_classes: int,
                   score_threshold: float, row_offset: int, rec: torch.Tensor, rec_count: torch.Tensor) -> bool:
    """One level of pqdet_head_conv_hits: appends the level's hit records to rec (B, cap, 6+C) / rec_count (B).
    -> False when the shape is outside the persistent tensor-core kernel (the caller takes the raw-head route)."""
    x = _req(x, "x")
    weight = _req(weight.reshape(weight.shape[0], -1), "weight")
    if bias is not None:
        bias = _req(bias, "bias")
    B, Cin, H, W = x.shape
    ch = 5 + num_classes
    if weight.shape[0] % ch or weight.shape[1] != Cin:
        raise ValueError("weight must be (A*(5+C), Cin)")
    A = weight.shape[0] // ch
    code = _lib.load().pqdet_head_conv_hits(_ptr(x), _ptr(weight), _ptr(bias), B, Cin, H, W, A, num_classes,
                                            float(score_threshold), int(row_offset), _ptr(rec), _ptr(rec_count),
                                            int(rec.shape[1]), _dev(x), _stream(x.device))
    if code == _lib.PQDET_ERR_UNSUPPORTED:
        return False
    _lib.check(code, "pqdet_head_conv_hits")
    return True


def records_nms(heads_t, keep_alive, rec: torch.Tensor, rec_count: torch.Tensor, max_det: int, want_index: bool,
                capacity: str = "compact"):
    """The fused kernel's back end on hit records (pqdet_records_nms).  -> det, idx, meta as decode_nms_fused."""
    device = rec.device
    B = heads_t.B
    det, idx, meta = alloc_fused_outputs(B, max_det, want_index, device)
    counts, ncand, status, work = meta[0:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:]
    _lib.check(_lib.load().pqdet_records_nms(ctypes.byref(heads_t), _ptr(rec), _ptr(rec_count), int(rec.shape[1]),
                                             _ptr(det), _ptr(idx), int(max_det), _ptr(counts), _ptr(ncand),
                                             _ptr(status), _ptr(work), 0, _lib.CAPACITY[capacity], _dev(rec),
                                             _stream(device)), "pqdet_records_nms")
    return det, idx, meta


def decode_nms_gather(heads_t, keep_alive, max_det: int, out, det_ptrs: Sequence[int], cnt_ptrs: Sequence[int], rank: int,
                      gather_cap: int, capacity: str = "compact", arr_ptrs: Optional[Sequence[int]] = None):
    """pqdet_decode_nms_gather: the fused kernel that also stores its rows / counts into every rank's gathered buffers
    (peer memory).  out = buffers of alloc_fused_outputs (reused across calls).  arr_ptrs: this rank's arrival counter
    in every rank's buffer (the kernel then signals per image; peer_wait on the receiving side).  -> det, meta."""
    raws, _ = keep_alive
    device = raws[0].device
    B = heads_t.B
    det, _, meta = out
    counts, ncand, status, work = meta[0:B], meta[B:2 * B], meta[2 * B:3 * B], meta[3 * B:]
    key = (work.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
    armed = 1 if key in _FUSED_ARMED else 0
    _FUSED_ARMED.discard(key)
    n = len(det_ptrs)
    VP = ctypes.c_void_p * n
    _lib.check(_lib.load().pqdet_decode_nms_gather(ctypes.byref(heads_t), _ptr(det), int(max_det), _ptr(counts),
                                                   _ptr(ncand), _ptr(status), VP(*det_ptrs), VP(*cnt_ptrs),
                                                   VP(*arr_ptrs) if arr_ptrs is not None else None, n, int(rank),
                                                   int(gather_cap), _ptr(work), armed, _lib.CAPACITY[capacity],
                                                   _dev(raws[0]), _stream(device)), "pqdet_decode_nms_gather")
    _FUSED_ARMED.add(key)
    return det, meta


def peer_wait(arrived: torch.Tensor, n: int, expected: int, err_flag: Optional[torch.Tensor] = None) -> None:
    """pqdet_peer_wait: block the stream (one warp, programmatic dependent) until the n arrival counters in `arrived`
    (uint32 bit patterns in an int32 CUDA tensor) have reached `expected` modulo 2^32."""
    device = arrived.device
    _lib.check(_lib.load().pqdet_peer_wait(_ptr(arrived), int(n), int(expected) & 0xffffffff,
                                           _ptr(err_flag) if err_flag is not None else None, _dev(arrived),
                                           _stream(device)), "pqdet_peer_wait")
