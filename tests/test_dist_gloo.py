"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing in pqdet_b200/dist.py: the 7-float loss
all-reduce equals the reference's mean of replica means, and gathered detections come back in image order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pqdet_b200 import dist as pqd
        total = 10
        lo, hi = pqd.shard_range(total, rank, world)
        g = torch.Generator().manual_seed(0)
        per_image = torch.rand((total, 7), generator=g)                 # per-image loss contributions
        local = per_image[lo:hi]
        losses = {k: local[:, i].mean().reshape(1) for i, k in enumerate(['loss', 'giou_loss', 'conf_loss', 'class_loss'])}
        losses['loss_per_branch'] = [local[:, 4 + i].mean().reshape(1) for i in range(3)]
        red = pqd.reduce_losses(losses, hi - lo)
        want = per_image.mean(dim=0)
        got = torch.cat([red['loss'], red['giou_loss'], red['conf_loss'], red['class_loss']] + red['loss_per_branch'])
        ok_loss = bool(torch.allclose(got, want, rtol=1e-6, atol=1e-7))
        # detections: image i has (i % 4) rows filled with the value i
        counts = torch.tensor([(i % 4) for i in range(lo, hi)], dtype=torch.int32)
        det = torch.zeros((hi - lo, 5, 6))
        for j, i in enumerate(range(lo, hi)):
            det[j, :i % 4] = float(i)
        flat = pqd.flatten_gathered(pqd.gather_detections(det, counts))
        ok_det = len(flat) == total and all(t.shape[0] == i % 4 and bool((t == float(i)).all()) for i, t in enumerate(flat))
        # the lean form: the kernel's own (4+5L,) result vector, global batch known -> one collective
        L = 3
        out_vec = torch.zeros((4 + 5 * L,))
        out_vec[0:4] = local[:, 0:4].mean(dim=0)
        out_vec[4 + 4 * L:4 + 5 * L] = local[:, 4:7].mean(dim=0)
        mine = {k: out_vec[i:i + 1] for i, k in enumerate(['loss', 'giou_loss', 'conf_loss', 'class_loss'])}
        mine['loss_per_branch'] = [out_vec[4 + 4 * L + i:5 + 4 * L + i] for i in range(L)]
        mine['loss'].pq_out = out_vec
        red2 = pqd.reduce_losses(mine, hi - lo, global_batch=total)
        got2 = torch.cat([red2['loss'], red2['giou_loss'], red2['conf_loss'], red2['class_loss']] + red2['loss_per_branch'])
        ok_loss = ok_loss and bool(torch.allclose(got2, want, rtol=1e-6, atol=1e-7))
        # fixed-capacity gather (equal shards: 5 images each): one collective, no host round trip
        all_det, all_counts, bufs = pqd.gather_detections_fixed(det, counts, 3)
        ok_fix = tuple(all_det.shape) == (world, hi - lo, 3, 6) and all_counts.dtype == torch.int32
        for r in range(world):
            rl, rh = pqd.shard_range(total, r, world)
            for j, i in enumerate(range(rl, rh)):
                n = min(i % 4, 3)
                ok_fix = ok_fix and int(all_counts[r, j]) == n and bool((all_det[r, j, :n] == float(i)).all())
        all_det2, _, _ = pqd.gather_detections_fixed(det, counts, 3, out=bufs)        # buffer reuse
        ok_det = ok_det and ok_fix and bool(torch.equal(all_det, all_det2))
        q.put((rank, ok_loss, ok_det))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_loss_allreduce_and_detection_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
