"""Multi-GPU check of the sharded path over NCCL (not collected by pytest: run under torchrun on an N-GPU box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multi_gpu_check.py

Every rank owns a contiguous block of images (pqdet_b200.dist.shard_range).  Checks, per rank:
  1. eval: the rank's fused decode+NMS rows are bit-identical to the rows the single-GPU run over the WHOLE batch
     produces for the same images; gather_detections returns them in global image order on every rank;
  2. train: reduce_losses (one all_reduce of 7 sums + the batch size) equals the whole-batch loss of one GPU.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pqdet_b200 import dist as pqd, fused, synth
    from pqdet_b200.interpreter import DetectionHead
    from pqdet_b200.train_dataset import LabelAssigner
    C, size, total = 20, 512, 50                       # 50 images: uneven shards for world = 4 or 8
    heads = synth.make_heads(total, C, size, "sparse", seed=3)
    orig = torch.tensor([[375., 500.], [float(size), float(size)]]).repeat(total // 2, 1)
    lo, hi = pqd.shard_range(total, rank, world)
    full = fused.decode_nms([h.to(dev) for h in heads], synth.FPN_STRIDES, C, (size, size), orig.to(dev), "voc", 0.1, 0.45)
    mine = fused.decode_nms([h[lo:hi].to(dev) for h in heads], synth.FPN_STRIDES, C, (size, size), orig[lo:hi].to(dev),
                            "voc", 0.1, 0.45)
    for j, b in enumerate(range(lo, hi)):
        assert torch.equal(mine[j], full[b]), "rank %d image %d differs from the single-GPU run" % (rank, b)
    flat = pqd.flatten_gathered(pqd.gather_detections(mine.det, mine.counts.to(dev)))
    assert len(flat) == total
    for b in range(total):
        assert torch.equal(flat[b], full[b]), "gathered image %d" % b
    # the same gather without a collective: rows stored into every rank's buffers through peer memory
    try:
        from pqdet_b200 import _ops
        Bs = total // world                                  # equal shards for the fixed-shape form
        lo2, hi2 = rank * Bs, (rank + 1) * Bs
        sub = [h[lo2:hi2].to(dev) for h in heads]
        hs, keep = _ops.make_heads(sub, synth.FPN_STRIDES, C, (size, size), orig[lo2:hi2].to(dev), "voc", 0.1, 0.45,
                                   "auto_cuda", "tv_cuda")
        for sync in ("signal", "barrier"):                   # arrival counters + wait kernel | barrier per step
            pg = pqd.PeerGather(Bs, 512, dev, sync=sync)
            for _ in range(3):                               # repeated calls reuse the buffers
                all_det, all_cnt, _, _ = pg.decode_nms(hs, keep)
            torch.cuda.synchronize()
            assert int(pg.err) == 0, "peer wait timed out on source rank %d" % (int(pg.err) - 1)
            for r in range(world):
                for j in range(Bs):
                    k = int(all_cnt[r, j])
                    assert torch.equal(all_det[r, j, :k], full[r * Bs + j]), "peer gather (%s) rank %d image %d" % (sync, r, j)
            dist.barrier()
        peer = "peer-memory gather ok (signal and barrier forms)"
    except RuntimeError as e:
        peer = "peer-memory gather unavailable: %s" % (str(e).splitlines()[0][:120],)
    # training: local loss over the shard, reduced, vs the whole batch on one GPU
    gts = synth.make_gt(total, C, size, 1, 12, seed=0)
    out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
    la = LabelAssigner(C, device=dev)
    raws = synth.make_train_heads(total, C, size, seed=0, device=dev)
    opts = [dict(classes=C, stride=s, bbox_loss="giou", ignore_thresh=0.5, l1_loss_gain=0.05) for s in synth.FPN_STRIDES]
    head = DetectionHead(opts)
    whole = head(raws, la.create_label_batch(gts, out_sizes))
    local_out = head([r[lo:hi].contiguous() for r in raws], la.create_label_batch(gts[lo:hi], out_sizes))
    red = pqd.reduce_losses(local_out, hi - lo)
    for k in ("loss", "giou_loss", "conf_loss", "class_loss"):
        a, b = float(red[k]), float(whole[k])
        assert abs(a - b) <= 1e-5 * abs(b), (k, a, b)
    for a, b in zip(red["loss_per_branch"], whole["loss_per_branch"]):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
    try:
        for sync in ("signal", "barrier"):                   # arrival counters + wait kernel | barrier per step
            pr = pqd.PeerReduce(dev, sync=sync)
            for _ in range(5):
                red2 = pr.reduce_losses(local_out, hi - lo, total)
            for k in ("loss", "giou_loss", "conf_loss", "class_loss"):
                assert abs(float(red2[k]) - float(whole[k])) <= 1e-5 * abs(float(whole[k])), (sync, k, float(red2[k]), float(whole[k]))
            for a, b in zip(red2["loss_per_branch"], whole["loss_per_branch"]):
                assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b))
            assert int(pr.err) == 0, "peer wait timed out on source rank %d" % (int(pr.err) - 1)
            dist.barrier()
        peer += ", peer-memory loss reduce ok (signal and barrier forms)"
    except RuntimeError as e:
        peer += ", peer-memory loss reduce unavailable: %s" % (str(e).splitlines()[0][:120],)
    dist.barrier()
    if rank == 0:
        print("multi_gpu_check ok: world=%d, %d images, sharded rows bit-identical, reduced loss %.6f == %.6f; %s"
              % (world, total, float(red["loss"]), float(whole["loss"]), peer))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
