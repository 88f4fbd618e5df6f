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

FUSED_MAX_DET = 2048          # == kCapM of csrc/nms.cu: a fused image never keeps more than its candidates


class Detections:
    """Result of decode_nms: padded device buffers + per-image views after one host read."""

    def __init__(self, det, idx, meta, B, spill=None):
        self.det, self.idx, self.meta, self.B = det, idx, meta, B
        self._spill = spill or {}
        self._host = None

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
        if b in self._spill:
            return self._spill[b][0]
        return self.det[b, :int(self.counts[b])]

    def indices(self, b: int) -> torch.Tensor:
        """row*C + class of every detection of image b (requires return_index=True)."""
        if b in self._spill:
            return self._spill[b][1]
        return self.idx[b, :int(self.counts[b])].to(torch.int64)

    def to_reference_list(self) -> List[torch.Tensor]:
        """What the reference's per-image loop produces: tensors (K,6), or shape (0,) when empty."""
        out = []
        for b in range(self.B):
            d = self[b]
            out.append(d if d.shape[0] else torch.tensor([]).to(d))
        return out


def decode_nms(heads: Sequence[torch.Tensor], strides: Sequence[int], num_classes: int, input_size,
               batch_original_size, dataset: str = "voc", score_threshold: float = 0.1,
               iou_threshold: float = 0.45, return_index: bool = False, resolve_overflow: bool = True,
               nms_mode: Optional[str] = None, iou_round: Optional[str] = None) -> Detections:
    """heads: raw [yolo] inputs (B, A*(5+C), H_l, W_l) in cfg order with their strides.
    input_size (h, w) -- pass a tuple/CPU tensor (a CUDA tensor costs a sync);
    batch_original_size (B,2) or (2,) (h, w).  dataset in {'voc','coco','visdrone'} picks the affine.
    With resolve_overflow (default) images whose candidates do not fit the on-chip lists are re-run
    through the general path, so the result is always complete (costs the host read of `status`)."""
    m, r = config.nms_modes()
    nms_mode, iou_round = nms_mode or m, iou_round or r
    h, keep = _ops.make_heads(heads, strides, num_classes, input_size, batch_original_size, dataset,
                              score_threshold, iou_threshold, nms_mode, iou_round)
    B = h.B
    det, idx, meta = _ops.decode_nms_fused(h, keep, FUSED_MAX_DET, return_index)
    res = Detections(det, idx, meta, B)
    if not resolve_overflow or B == 0:
        return res
    status = res.host_meta()[2]
    over = torch.nonzero(status & _lib.ST_CAND_OVERFLOW).reshape(-1)
    if over.numel() == 0:
        return res
    ids = over.to(torch.int32).to(det.device)
    n_sel = int(ids.numel())
    cap, max_det = max(n_sel * 32768, 1 << 16), 8192
    for _ in range(3):
        gdet, gidx, gmeta, needed = _ops.nms_general(heads_t=h, keep_alive=keep, image_ids=ids, n_images=n_sel,
                                                     max_det=max_det, cand_capacity=cap,
                                                     want_index=return_index, out_by_position=True)
        host = torch.cat([gmeta.to(torch.int64), needed]).cpu()
        gcounts, gncand, gstatus = host[0:n_sel], host[n_sel:2 * n_sel], host[2 * n_sel:3 * n_sel]
        if bool((gstatus & _lib.ST_CAND_OVERFLOW).any()):
            cap = max(int(host[3 * n_sel]), cap * 2)
            continue
        if bool((gstatus & _lib.ST_DET_TRUNCATED).any()):
            max_det = int(gcounts.max())
            continue
        hm = res.host_meta()
        for i, b in enumerate(over.tolist()):
            k = int(gcounts[i])
            res._spill[b] = (gdet[i, :k], gidx[i, :k].to(torch.int64) if return_index else None)
            hm[0, b], hm[1, b], hm[2, b] = k, int(gncand[i]), 0
        return res
    raise _lib.PqdetError("general NMS path did not converge on a workspace size")
