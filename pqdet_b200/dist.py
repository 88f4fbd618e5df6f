"""Multi-GPU plumbing for the hot path: one process per GPU, shard by image, no data-path collective.

The reference is single-process nn.DataParallel (tools.py:215-216): it scatters inputs + dense labels
over the GPUs every step, gathers the (B,N,5+C) predictions to GPU 0 and averages per-replica losses
(model/loss.py:105-108, trainer.py:233).  Here every rank owns a contiguous block of images end to
end; the only traffic (NCCL over NVLink) is

  * training: ONE all_reduce(SUM) of 7 floats per step -- [loss, bbox, conf, cls, branch0..2] as
    per-rank sums over images -- then a divide by the global batch, which equals the reference's
    mean of replica means when shards are equal;
  * eval: all_gather of the per-image counts + of the padded (K_max, 6) detection rows
    (a few KB per image), or nothing at all if every rank consumes its own shard.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of images for `rank`; the first (total % world) ranks get one extra."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_losses(losses: dict, local_batch: int, group=None) -> dict:
    """losses: the dict DetectionHead.forward returns on this rank (batch means over local_batch).
    -> the same dict with every entry replaced by the global-batch mean.  One 7-float all_reduce."""
    keys = ['loss', 'giou_loss', 'conf_loss', 'class_loss']
    parts = [losses[k].reshape(-1)[:1] for k in keys] + [b.reshape(-1)[:1] for b in losses['loss_per_branch']]
    nb = len(losses['loss_per_branch'])
    vec = torch.cat(parts).detach().to(torch.float32) * float(local_batch)
    vec = torch.cat([vec, vec.new_tensor([float(local_batch)])])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    vec = vec[:-1] / vec[-1]
    out = {k: vec[i:i + 1] for i, k in enumerate(keys)}
    out['loss_per_branch'] = [vec[4 + i:5 + i] for i in range(nb)]
    return out


def gather_detections(det: torch.Tensor, counts: torch.Tensor, group=None):
    """det (B_local, K_cap, 6) padded rows, counts (B_local) int -> on every rank:
    list over ranks of (det_r trimmed to that rank's max K, counts_r), in rank (= image) order.
    Shards may have different B_local."""
    if not (dist.is_available() and dist.is_initialized()):
        k = int(counts.max()) if counts.numel() else 0
        return [(det[:, :k], counts)]
    world = dist.get_world_size(group)
    dev = det.device
    meta = torch.tensor([det.shape[0], int(counts.max()) if counts.numel() else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    b_max = max(int(m[0]) for m in metas)
    k_max = max(int(m[1]) for m in metas)
    pad_counts = torch.zeros((b_max,), dtype=torch.int32, device=dev)
    pad_counts[:det.shape[0]] = counts.to(torch.int32)
    pad_det = torch.zeros((b_max, max(k_max, 1), 6), dtype=torch.float32, device=dev)
    kk = min(k_max, det.shape[1])
    pad_det[:det.shape[0], :kk] = det[:, :kk]
    all_counts = [torch.zeros_like(pad_counts) for _ in range(world)]
    all_det = [torch.zeros_like(pad_det) for _ in range(world)]
    dist.all_gather(all_counts, pad_counts, group=group)
    dist.all_gather(all_det, pad_det, group=group)
    out = []
    for r in range(world):
        b_r, k_r = int(metas[r][0]), int(metas[r][1])
        out.append((all_det[r][:b_r, :k_r], all_counts[r][:b_r]))
    return out


def flatten_gathered(gathered) -> List[torch.Tensor]:
    """-> one (K,6) tensor per image in global image order."""
    res = []
    for det, counts in gathered:
        for b in range(det.shape[0]):
            res.append(det[b, :int(counts[b])])
    return res
