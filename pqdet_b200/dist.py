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


def reduce_losses(losses: dict, local_batch: int, group=None, global_batch: Optional[int] = None) -> dict:
    """losses: the dict DetectionHead.forward returns on this rank (batch means over local_batch).
    -> the same dict with every entry replaced by the global-batch mean (the reference's mean of replica means,
    model/loss.py:105-108 + trainer.py:233, for any shard sizes).

    With `global_batch` given (the trainer knows it) this is ONE collective on the kernel's own output vector and at
    most one scaling kernel: all_reduce(AVG) when the shards are equal, else the vector is pre-scaled by
    local_batch / global_batch and summed.  Without it the batch size travels with the sums (7 + 1 floats)."""
    keys = ['loss', 'giou_loss', 'conf_loss', 'class_loss']
    nb = len(losses['loss_per_branch'])
    live = dist.is_available() and dist.is_initialized()
    raw = getattr(losses['loss'], 'pq_out', None)          # DetectionHead's (4+5L,) result vector, if this is ours
    if global_batch is not None and raw is not None:
        vec = raw.detach().clone()
        world = dist.get_world_size(group) if live else 1
        if local_batch * world == global_batch and (not live or dist.get_backend(group) == "nccl"):
            if live:
                dist.all_reduce(vec, op=dist.ReduceOp.AVG, group=group)          # NCCL-only reduction op
        else:
            vec.mul_(float(local_batch) / float(global_batch))
            if live:
                dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        L = nb
        out = {k: vec[i:i + 1] for i, k in enumerate(keys)}
        out['loss_per_branch'] = [vec[4 + 4 * L + i:5 + 4 * L + i] for i in range(L)]
        return out
    parts = [losses[k].reshape(-1)[:1] for k in keys] + [b.reshape(-1)[:1] for b in losses['loss_per_branch']]
    vec = torch.cat(parts).detach().to(torch.float32) * float(local_batch)
    vec = torch.cat([vec, vec.new_tensor([float(local_batch)])])
    if live:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    vec = vec[:-1] / vec[-1]
    out = {k: vec[i:i + 1] for i, k in enumerate(keys)}
    out['loss_per_branch'] = [vec[4 + i:5 + i] for i in range(nb)]
    return out


def gather_detections_fixed(det: torch.Tensor, counts: torch.Tensor, k_cap: int, group=None, out=None):
    """Eval gather without a host round trip: every rank contributes the same shape, so ONE all_gather moves
    everything.  det (B, K, 6) padded rows, counts (B) int32 (device) -> (all_det (world, B, k_cap, 6),
    all_counts (world, B) int32) on every rank, rank-major = global image order for equal contiguous shards.
    Rows beyond k_cap are dropped (counts are clamped); the packed record per image is [count | k_cap rows], the count
    bit-cast into the float buffer.  `out` = a previous result's packed buffers to reuse."""
    B = det.shape[0]
    k = min(k_cap, det.shape[1])
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rec = 1 + k_cap * 6
    if out is None:
        send = torch.zeros((B, rec), dtype=torch.float32, device=det.device)
        recv = torch.empty((world, B, rec), dtype=torch.float32, device=det.device)
    else:
        send, recv = out
    send[:, 0] = counts.to(torch.int32).clamp(max=k_cap).view(torch.float32)
    send[:, 1:1 + k * 6] = det[:, :k].reshape(B, k * 6)
    if world > 1:
        dist.all_gather_into_tensor(recv.view(world * B, rec), send, group=group)
    else:
        recv[0].copy_(send)
    all_counts = recv[:, :, 0].contiguous().view(torch.int32)
    all_det = recv[:, :, 1:].view(world, B, k_cap, 6)
    return all_det, all_counts, (send, recv)


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


class PeerGather:
    """Eval gather over peer memory instead of a collective (one node, NVLink / NVSwitch): every rank's fused
    decode+NMS launch stores its kept rows and counts straight into the gathered buffers of ALL ranks
    (pqdet_decode_nms_gather; the buffers are torch symmetric memory, mapped into every process), then a device-side
    barrier on the symmetric-memory signal pads.  No data-path collective, no host round trip.

        pg = PeerGather(B_local, k_cap, device)            # once (collective: rendezvous)
        all_det, all_counts = pg.decode_nms(heads_t, keep)  # every step; (world, B, k_cap, 6), (world, B) int32

    Raises RuntimeError where symmetric memory is unavailable (callers fall back to gather_detections_fixed)."""

    def __init__(self, B: int, k_cap: int, device, group=None, sync: str = "signal"):
        """sync = 'signal': the kernel bumps an arrival counter on every peer per image and a one-warp wait kernel on
        the receiving side replaces the barrier (consecutive steps keep overlapping); 'barrier': a symmetric-memory
        barrier after every launch."""
        import torch.distributed._symmetric_memory as symm
        if sync not in ("signal", "barrier"):
            raise ValueError("sync must be 'signal' or 'barrier'")
        self.sync = sync
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerGather maps at most 8 peers (one node)")
        self.B, self.k_cap, self.device = B, k_cap, torch.device(device)
        n_det = self.world * B * k_cap * 6
        # one symmetric allocation: [gathered rows | gathered counts (int32 bit patterns) | arrival counters (8)]
        n_cnt = self.world * B
        self.buf = symm.empty(n_det + n_cnt + 8, dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.all_det = self.buf[:n_det].view(self.world, B, k_cap, 6)
        self.all_counts = self.buf[n_det:n_det + n_cnt].view(torch.int32).view(self.world, B)
        self.arrived = self.buf[n_det + n_cnt:].view(torch.int32)      # [source rank]: images of that rank landed here
        self.all_counts.zero_()
        self.arrived.zero_()
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._det_ptrs = ptrs
        self._cnt_ptrs = [p + n_det * 4 for p in ptrs]
        # this rank's counter inside every rank's buffer
        self._arr_ptrs = [p + (n_det + n_cnt + self.rank) * 4 for p in ptrs]
        self.err = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._steps = 0
        self._out = None
        torch.cuda.synchronize(self.device)
        self.hdl.barrier()

    def decode_nms(self, heads_t, keep_alive, max_det: int = 2048, capacity: str = "compact"):
        """Fused decode+NMS of this rank's images; returns the gathered (all_det, all_counts) views, complete on every
        rank once the call's stream work is done, and the local (det, meta) buffers."""
        from . import _ops
        if self._out is None:
            self._out = _ops.alloc_fused_outputs(self.B, max_det, False, self.device)
        if self.sync == "barrier":
            det, meta = _ops.decode_nms_gather(heads_t, keep_alive, max_det, self._out, self._det_ptrs, self._cnt_ptrs,
                                               self.rank, self.k_cap, capacity)
            self.hdl.barrier()                                # every rank's rows have landed everywhere
            return self.all_det, self.all_counts, det, meta
        det, meta = _ops.decode_nms_gather(heads_t, keep_alive, max_det, self._out, self._det_ptrs, self._cnt_ptrs,
                                           self.rank, self.k_cap, capacity, arr_ptrs=self._arr_ptrs)
        self._steps += 1
        # every source rank has delivered B images per step (equal shards: B is the same on every rank)
        _ops.peer_wait(self.arrived, self.world, self._steps * self.B, self.err)
        return self.all_det, self.all_counts, det, meta


class PeerReduce:
    """reduce_losses over peer memory instead of a collective: every rank stores its scaled (4+5L)-float loss vector
    into its row of every rank's symmetric buffer (pqdet_peer_publish), a device-side barrier, a local sum over the rows
    in rank order (pqdet_peer_sum_rows).  Deterministic, no host round trip; the fixed cost of the NCCL all_reduce
    (~80 us through torch for 19 floats) drops to two tiny launches and a signal-pad barrier."""

    ROW = 64

    def __init__(self, device, group=None, sync: str = "barrier"):
        """sync = 'barrier' (default): a symmetric-memory barrier between publish and sum; 'signal': arrival counters
        and pqdet_peer_wait as in PeerGather (same results; its step time was not reproducible between boxes, see
        DESIGN section 7, so it is not the default here)."""
        import torch.distributed._symmetric_memory as symm
        if sync not in ("signal", "barrier"):
            raise ValueError("sync must be 'signal' or 'barrier'")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerReduce maps at most 8 peers (one node)")
        self.device = torch.device(device)
        # two row sets used alternately: a fast rank's step k+1 never touches rows a slow rank still sums for step k,
        # and by the time it reaches step k+2 it has passed step k+1's barrier, which the slow rank only enters after
        # its step-k sum (stream order) - one barrier per step suffices
        # [row set 0 | row set 1 | arrival counters (8)]: a publish adds 1 to this rank's counter in every rank's buffer,
        # pqdet_peer_wait on the receiving side replaces the barrier kernel (sync='barrier' keeps it)
        n_rows = 2 * self.world * self.ROW
        self.buf = symm.empty(n_rows + 8, dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.buf.zero_()
        self._ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.arrived = self.buf[n_rows:].view(torch.int32)
        self._arr_ptrs = [p + (n_rows + self.rank) * 4 for p in self._ptrs]
        self.err = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.sync = sync
        self._step = 0
        torch.cuda.synchronize(self.device)
        self.hdl.barrier()

    def reduce_losses(self, losses: dict, local_batch: int, global_batch: int) -> dict:
        """Same result as dist.reduce_losses(losses, local_batch, global_batch=...): the global-batch means."""
        import ctypes
        from . import _lib
        raw = getattr(losses['loss'], 'pq_out', None)
        if raw is None:
            raise ValueError("PeerReduce needs the loss dict DetectionHead returns (its result vector)")
        n = raw.numel()
        nb = len(losses['loss_per_branch'])
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        VP = ctypes.c_void_p * self.world
        lib = _lib.load()
        half = (self._step & 1) * self.world * self.ROW * 4          # byte offset of this step's row set
        self._step += 1
        signal = self.sync == "signal"
        _lib.check(lib.pqdet_peer_publish(ctypes.c_void_p(raw.data_ptr()), n, float(local_batch) / float(global_batch),
                                          VP(*[p + half for p in self._ptrs]),
                                          VP(*self._arr_ptrs) if signal else None, self.world, self.rank, self.ROW,
                                          dev_index, st), "pqdet_peer_publish")
        if signal:
            from . import _ops
            _ops.peer_wait(self.arrived, self.world, self._step, self.err)      # every rank has published this step
        else:
            self.hdl.barrier()
        vec = torch.empty((n,), dtype=torch.float32, device=self.device)
        _lib.check(lib.pqdet_peer_sum_rows(ctypes.c_void_p(self.buf.data_ptr() + half), n, self.world, self.ROW,
                                           ctypes.c_void_p(vec.data_ptr()), dev_index, st), "pqdet_peer_sum_rows")
        keys = ['loss', 'giou_loss', 'conf_loss', 'class_loss']
        out = {k: vec[i:i + 1] for i, k in enumerate(keys)}
        out['loss_per_branch'] = [vec[4 + 4 * nb + i:5 + 4 * nb + i] for i in range(nb)]
        return out
