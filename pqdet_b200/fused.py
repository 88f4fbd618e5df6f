"""The fused eval path: raw heads -> detections in one kernel launch.

Replaces, for a whole batch, the reference sequence
    Decode x L -> view/cat -> recover_bboxes_prediction_* -> for each image: tools.torch_nms
(predict.py:33-45, eval/evaluator.py:48-59, test.py:178-185) with results identical to running
tools.torch_nms on our own decoded + recovered boxes (keep lists bit-exact).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, _ops, config

FUSED_MAX_DET = 2048          # the larger capacity class' candidate list: a fused image never keeps more than that


class Detections:
    """Result of decode_nms: padded device buffers + per-image views after one host read."""

    def __init__(self, det, idx, meta, B):
        self.det, self.idx, self.meta, self.B = det, idx, meta, B
        self._spill_batch = None          # (det, idx, host meta, {image id: position}) of the general path
        self.general_first = False        # the whole batch went straight to the general path
        self._host = None

    @property
    def _spill(self):
        """image ids that were resolved by the general path (overflowed the fused kernel)."""
        return self._spill_batch[3] if self._spill_batch else {}

    @property
    def images_via_general_path(self) -> int:
        return self.B if self.general_first else len(self._spill)

    def host_meta(self) -> torch.Tensor:
        """(3,B) int32 on the host: counts, ncand, status.  The only device->host sync."""
        if self._host is None:
            self._host = self.meta[:3 * self.B].view(3, self.B).cpu()
        return self._host

    @property
    def counts(self) -> torch.Tensor:
        return self.host_meta()[0]

    def __len__(self):
        return self.B

    def __getitem__(self, b: int) -> torch.Tensor:
        """(K_b, 6) rows [x1,y1,x2,y2,score,class] of image b, descending score."""
        if self._spill_batch and b in self._spill_batch[3]:
            i = self._spill_batch[3][b]
            return self._spill_batch[0][i, :int(self._spill_batch[2][0, i])]
        return self.det[b, :int(self.counts[b])]

    def indices(self, b: int) -> torch.Tensor:
        """row*C + class of every detection of image b (requires return_index=True)."""
        if self._spill_batch and b in self._spill_batch[3]:
            i = self._spill_batch[3][b]
            return self._spill_batch[1][i, :int(self._spill_batch[2][0, i])].to(torch.int64)
        return self.idx[b, :int(self.counts[b])].to(torch.int64)

    def to_numpy_list(self):
        """Evaluator ingest (eval/evaluator.py:53-61 without the per-image loop of device syncs): ONE
        device->host copy of the padded rows, then per-image numpy views [(K_b, 6) float32]."""
        counts = self.counts
        kmax = int(counts.max()) if self.B else 0
        host = self.det[:, :kmax].contiguous().cpu().numpy() if kmax else None
        out = []
        for b in range(self.B):
            if self._spill_batch and b in self._spill_batch[3]:
                out.append(self[b].cpu().numpy())
            else:
                k = int(counts[b])
                out.append(host[b, :k] if k else __import__("numpy").zeros((0, 6), "float32"))
        return out

    def to_reference_list(self) -> List[torch.Tensor]:
        """What the reference's per-image loop produces: tensors (K,6), or shape (0,) when empty."""
        out = []
        for b in range(self.B):
            d = self[b]
            out.append(d if d.shape[0] else torch.tensor([]).to(d))
        return out


class StrategyHints(dict):
    """Caller-held memory for strategy='auto': workload signature -> 'compact' | 'large' | 'general', the route the
    last call with that signature suggests for the next one (compact / large = the fused kernel's capacity classes,
    general = the bucketed path for dense scenes).  The library keeps no state of its own: pass one of these as
    `hints=` from whatever owns the evaluation loop (the evaluator hook of install.py holds one per Evaluator);
    without it every call starts on the compact fused kernel."""


_SLOTS = {}


def _large_class_slots(device) -> int:
    """Images the large capacity class holds in one resident wave (4 CTAs per SM)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SLOTS:
        _SLOTS[idx] = 4 * torch.cuda.get_device_properties(idx).multi_processor_count
    return _SLOTS[idx]


def _next_route(ncand_max: int, ncand_mean: float) -> str:
    """Route for a workload whose images have this many candidates (what a complete run reports)."""
    _, m_small = _lib.CAPACITY_LIMITS["compact"]
    _, m_large = _lib.CAPACITY_LIMITS["large"]
    if ncand_max <= m_small:
        return "compact"
    if ncand_max <= m_large and ncand_mean <= m_large / 2:
        return "large"
    return "general"


def _general(h, keep, ids, n_sel, return_index, by_position, cap, max_det):
    """Run the general path until its workspace / output capacities fit (normally one attempt)."""
    for _ in range(4):
        gdet, gidx, gmeta, needed = _ops.nms_general(heads_t=h, keep_alive=keep, image_ids=ids, n_images=n_sel,
                                                     max_det=max_det, cand_capacity=cap, want_index=return_index,
                                                     out_by_position=by_position)
        host = torch.cat([gmeta.to(torch.int64), needed]).cpu()           # the one device->host read
        R = n_sel if by_position else h.B
        gcounts, gncand, gstatus = host[0:R], host[R:2 * R], host[2 * R:3 * R]
        if bool((gstatus & _lib.ST_CAND_OVERFLOW).any()):
            cap = max(int(host[3 * R]), cap * 2)
            continue
        if bool((gstatus & _lib.ST_DET_TRUNCATED).any()):
            max_det = int(gcounts.max())
            continue
        return gdet, gidx, gmeta, host[:3 * R].view(3, R).to(torch.int32)
    raise _lib.PqdetError("general NMS path did not converge on a workspace size")


def decode_nms(heads: Sequence[torch.Tensor], strides: Sequence[int], num_classes: int, input_size,
               batch_original_size, dataset: str = "voc", score_threshold: float = 0.1,
               iou_threshold: float = 0.45, return_index: bool = False, resolve_overflow: bool = True,
               nms_mode: Optional[str] = None, iou_round: Optional[str] = None,
               strategy: str = "auto", hints: Optional[StrategyHints] = None) -> Detections:
    """heads: raw [yolo] inputs (B, A*(5+C), H_l, W_l) in cfg order with their strides.
    input_size (h, w) -- pass a tuple/CPU tensor (a CUDA tensor costs a sync);
    batch_original_size (B,2) or (2,) (h, w).  dataset in {'voc','coco','visdrone'} picks the affine.
    With resolve_overflow (default) images whose candidates do not fit the on-chip lists are re-run
    through the general path, so the result is always complete (costs the host read of `status`).
    strategy: 'compact' / 'large' (the one-launch kernel with that capacity class + per-image fallback; 'fused' =
    'compact'), 'general' (bucketed global-memory path for every image: the right choice for dense scenes such as
    VisDrone), or 'auto' (compact first; with a caller-held `hints` object it remembers per workload signature
    what the last call suggested - the large lists, or straight to the general path when most images overflowed)."""
    m, r = config.nms_modes()
    nms_mode, iou_round = nms_mode or m, iou_round or r
    h, keep = _ops.make_heads(heads, strides, num_classes, input_size, batch_original_size, dataset,
                              score_threshold, iou_threshold, nms_mode, iou_round)
    B = h.B
    if B == 0:                                       # empty batch: nothing to launch
        dev = heads[0].device
        return Detections(torch.empty((0, FUSED_MAX_DET, 6), device=dev),
                          torch.empty((0, FUSED_MAX_DET), dtype=torch.int32, device=dev) if return_index else None,
                          torch.zeros((2,), dtype=torch.int32, device=dev), 0)
    sig = (tuple(tuple(t.shape[1:]) for t in heads), num_classes, float(score_threshold), dataset)
    route = hints.get(sig, "compact") if (strategy == "auto" and hints is not None) else "compact"
    if route == "compact" and strategy == "auto" and B <= _large_class_slots(heads[0].device):
        # every image gets a CTA of its own in either class: the 256-thread CTAs of the large class finish an image
        # sooner (one image: 46 against 58 us; 592 images: 82 against 96 us), and their lists are twice as long
        route = "large"
    if strategy in ("compact", "large"):
        route = strategy
    if score_threshold < 0:
        route = "general"                     # the fused kernels' keys order non-negative scores only
    if B and (strategy == "general" or route == "general"):
        gdet, gidx, gmeta, hm = _general(h, keep, None, B, return_index, False, max(B * 32768, 1 << 16), 8192)
        res = Detections(gdet, gidx, gmeta, B)
        res._host = hm
        res.general_first = True
        if hints is not None:
            hints[sig] = _next_route(int(hm[1].max()), float(hm[1].float().mean()))
        return res
    det, idx, meta = _ops.decode_nms_fused(h, keep, FUSED_MAX_DET, return_index, capacity=route)
    res = Detections(det, idx, meta, B)
    res.capacity = route
    if not resolve_overflow or B == 0:
        return res
    status = res.host_meta()[2]
    over = torch.nonzero(status & _lib.ST_CAND_OVERFLOW).reshape(-1)
    if hints is not None:
        if over.numel() == 0:
            # what the workload NEEDS (not what this batch size made convenient): the next batch may be larger
            nc = res.host_meta()[1]
            hints[sig] = _next_route(int(nc.max()), float(nc.float().mean()))
        elif route == "compact" and over.numel() * 2 <= B:
            hints[sig] = "large"                  # some images outgrew the compact lists: try the large ones next
        else:
            hints[sig] = "general" if over.numel() * 2 > B else route
    if over.numel() == 0:
        return res
    ids = over.to(torch.int32).to(det.device)
    n_sel = int(ids.numel())
    gdet, gidx, _, ghm = _general(h, keep, ids, n_sel, return_index, True, max(n_sel * 32768, 1 << 16), 8192)
    hm = res.host_meta()
    res._spill_batch = (gdet, gidx, ghm, {int(b): i for i, b in enumerate(over.tolist())})
    for i, b in enumerate(over.tolist()):
        hm[0, b], hm[1, b], hm[2, b] = ghm[0, i], ghm[1, i], 0
    return res


def features_nms(features: Sequence[torch.Tensor], weights: Sequence[torch.Tensor], biases, strides: Sequence[int],
                 num_classes: int, input_size, batch_original_size, dataset: str = "voc",
                 score_threshold: float = 0.1, iou_threshold: float = 0.45, return_index: bool = False,
                 nms_mode: Optional[str] = None, iou_round: Optional[str] = None, capacity: str = "compact",
                 hints: Optional[StrategyHints] = None) -> Detections:
    """The eval path starting one layer earlier: features[l] = the INPUT of level l's 1x1 head
    convolution, weights / biases = its parameters -> detections, identical to decode_nms on the raw heads that
    convolution produces.  The convolution runs on the tensor cores (TF32, like PyTorch's default) and its epilogue
    keeps only the rows whose objectness can pass the threshold (a few hundred of the 16 128 rows of a VOC image), so
    neither the raw heads nor the decoded prediction are ever written: the bound is reading the features.
    Levels whose shape the persistent kernel cannot take (H*W not a multiple of 128) make the whole call take the
    raw-head route: head convolution with the raw output materialised, then decode_nms."""
    m, r = config.nms_modes()
    nms_mode, iou_round = nms_mode or m, iou_round or r
    C = num_classes
    B = features[0].shape[0]
    dev = features[0].device
    biases = list(biases) if biases is not None else [None] * len(features)
    A = weights[0].shape[0] // (5 + C)

    def raw_route():
        raws = [_ops.head_conv_decode(f, w, bi, C, float(s), want_raw=True, want_decoded=False)
                for f, w, bi, s in zip(features, weights, biases, strides)]
        return decode_nms(raws, strides, C, input_size, batch_original_size, dataset, score_threshold, iou_threshold,
                          return_index, True, nms_mode, iou_round, hints=hints)
    if B == 0 or score_threshold < 0:
        return raw_route()
    cap_h = _lib.CAPACITY_LIMITS[capacity][0]
    rec = torch.empty((B, cap_h, 6 + C), dtype=torch.float32, device=dev)
    rec_count = torch.zeros((B,), dtype=torch.int32, device=dev)
    row_off = 0
    for f, w, bi in zip(features, weights, biases):
        if not _ops.head_conv_hits(f, w, bi, C, score_threshold, row_off, rec, rec_count):
            return raw_route()
        row_off += f.shape[2] * f.shape[3] * A
    h, keep = _ops.make_geometry(B, A, [tuple(f.shape[2:]) for f in features], strides, C, input_size,
                                 batch_original_size, dataset, score_threshold, iou_threshold, nms_mode, iou_round, dev)
    det, idx, meta = _ops.records_nms(h, keep, rec, rec_count, FUSED_MAX_DET, return_index, capacity)
    res = Detections(det, idx, meta, B)
    res.capacity = capacity
    over = torch.nonzero(res.host_meta()[2] & _lib.ST_CAND_OVERFLOW).reshape(-1)
    if over.numel():
        # dense images: materialise the raw heads of just those and resolve them through the general path
        sub = over.to(dev)
        raws = [_ops.head_conv_decode(f[sub].contiguous(), w, bi, C, float(s), want_raw=True, want_decoded=False)
                for f, w, bi, s in zip(features, weights, biases, strides)]
        orig = batch_original_size if isinstance(batch_original_size, torch.Tensor) else torch.tensor(batch_original_size)
        orig = orig.to(device=dev, dtype=torch.float32)
        o_sub = orig[sub] if orig.dim() == 2 else orig
        d = decode_nms(raws, strides, C, input_size, o_sub, dataset, score_threshold, iou_threshold, return_index,
                       True, nms_mode, iou_round, strategy="general")
        hm = res.host_meta()
        ghm = d.host_meta()
        res._spill_batch = (d.det, d.idx, ghm, {int(b): i for i, b in enumerate(over.tolist())})
        for i, b in enumerate(over.tolist()):
            hm[0, b], hm[1, b], hm[2, b] = ghm[0, i], ghm[1, i], 0
    return res


class HostDetections:
    """Result of decode_nms_host: everything already sits in pinned host memory (numpy views, no copies)."""

    def __init__(self, det, idx, meta, B, event):
        self.det, self.idx, self.meta, self.B, self._event = det, idx, meta, B, event

    def wait(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self

    @property
    def counts(self):
        return self.wait().meta[:self.B]

    @property
    def status(self):
        return self.wait().meta[2 * self.B:3 * self.B]

    def __len__(self):
        return self.B

    def __getitem__(self, b: int) -> torch.Tensor:
        return self.wait().det[b, :int(self.meta[b])]

    def to_numpy_list(self):
        self.wait()
        det, cnt = self.det.numpy(), self.meta.numpy()
        return [det[b, :cnt[b]] for b in range(self.B)]


def decode_nms_host(heads: Sequence[torch.Tensor], strides: Sequence[int], num_classes: int, input_size,
                    batch_original_size, dataset: str = "voc", score_threshold: float = 0.1,
                    iou_threshold: float = 0.45, return_index: bool = False, device=None, out=None,
                    nms_mode: Optional[str] = None, iou_round: Optional[str] = None,
                    capacity: str = "large") -> HostDetections:
    """decode_nms for head tensors in PINNED HOST memory; the detections come back in pinned host memory.
    (Default capacity class "large": with the heads behind PCIe the kernel is bound by the reads it has in flight,
    and 256-thread CTAs keep more of them in flight - 242 k against 225 k images/s on the headline workload.)
    The GPU reads only what the kernel touches (objectness planes + the channels of rows above threshold) straight
    over PCIe and writes the rows back, so there is no staging copy of the heads in either direction.  Images that
    overflow the fused kernel's on-chip lists are re-run from a device copy of just those images through the
    general path.  `out` = buffers of _ops.alloc_host_outputs to reuse across calls."""
    m, r = config.nms_modes()
    nms_mode, iou_round = nms_mode or m, iou_round or r
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h, keep = _ops.make_heads_host(heads, strides, num_classes, input_size, batch_original_size, dataset,
                                   score_threshold, iou_threshold, nms_mode, iou_round)
    B = h.B
    det, idx, meta, work = _ops.decode_nms_host(h, keep, FUSED_MAX_DET, return_index, device, out=out, capacity=capacity)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    res = HostDetections(det, idx, meta, B, ev)
    res._keep = (keep, work)
    if B == 0:
        return res
    status = res.status
    over = torch.nonzero(status & _lib.ST_CAND_OVERFLOW).reshape(-1)
    if over.numel():
        # dense images: stage just those on the device and use the general path; rows land in the host buffers
        sub = [t[over].to(device, non_blocking=True) for t in heads]
        orig = batch_original_size if isinstance(batch_original_size, torch.Tensor) else torch.tensor(batch_original_size)
        orig = orig.to(torch.float32)
        o_sub = orig[over] if orig.dim() == 2 else orig
        d = decode_nms(sub, strides, num_classes, input_size, o_sub.to(device), dataset, score_threshold, iou_threshold,
                       return_index, True, nms_mode, iou_round, strategy="general")
        hm = d.host_meta()
        grow = int(hm[0].max())
        if grow > res.det.shape[1]:
            ndet = torch.empty((B, grow, 6), dtype=torch.float32, pin_memory=True)
            ndet[:, :res.det.shape[1]] = res.det
            res.det = ndet
            if return_index:
                nidx = torch.empty((B, grow), dtype=torch.int32, pin_memory=True)
                nidx[:, :res.idx.shape[1]] = res.idx
                res.idx = nidx
        for i, b in enumerate(over.tolist()):
            k = int(hm[0, i])
            res.det[b, :k] = d[i].cpu()
            if return_index:
                res.idx[b, :k] = d.indices(i).to(torch.int32).cpu()
            res.meta[b], res.meta[B + b], res.meta[2 * B + b] = k, int(hm[1, i]), 0
    return res
