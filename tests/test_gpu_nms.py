"""-m gpu: NMS drop-in and the fused decode+NMS kernel vs the oracle, golden fixtures and (when
importable) the installed torchvision on both devices.  Keep lists are compared bit-exactly."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_close
from gpu_util import assert_same_detections, cuda, eval_chain_oracle
from oracle import pqdet_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_semantics():
    from pqdet_b200 import config
    old = config.nms_semantics
    yield
    config.nms_semantics = old


def test_torch_nms_golden_cpu_semantics():
    """tests/golden/nms.npz was produced by the reference on CPU tensors (torchvision CPU kernel)."""
    from pqdet_b200 import config, tools
    config.nms_semantics = "cpu"
    g = load_golden("nms")
    for name in sorted(k[:-3] for k in g if k.endswith("_in")):
        out = tools.torch_nms(cuda(g[name + "_in"]), float(g[name + "_thr"]), float(g[name + "_iou"]))
        want = g[name + "_out"]
        if want.size == 0:
            assert tuple(out.shape) == (0,), name            # tools.py:559-561
            continue
        assert_same_detections(out.cpu().numpy(), want, ties_unordered=(name == "vanilla_cpu"), what=name)


@pytest.mark.parametrize("sem", ["cuda", "cpu"])
@pytest.mark.parametrize("profile,C,size,kind,iou", [("sparse", 20, 512, "voc", 0.45),
                                                      ("dense", 10, 608, "visdrone", 0.45),
                                                      ("coco", 80, 608, "coco", 0.65)])
def test_torch_nms_dropin_vs_oracle(sem, profile, C, size, kind, iou):
    from pqdet_b200 import config, synth, tools
    config.nms_semantics = sem
    B = 2
    heads = synth.make_heads(B, C, size, profile, seed=11)
    pred = po.detect([h.numpy() for h in heads], C, synth.FPN_STRIDES)
    orig = np.array([[375., 500.], [float(size), float(size)]], np.float32)
    rec = po.recover(pred, (size, size), orig, kind)
    outs, idxs = tools.batched_torch_nms(cuda(rec), 0.1, iou, return_index=True)
    g_outs, g_idxs = tools.batched_torch_nms(cuda(rec), 0.1, iou, return_index=True, strategy="general")
    for b in range(B):
        assert torch.equal(outs[b], g_outs[b]) and torch.equal(idxs[b], g_idxs[b])    # fused == general path
    for b in range(B):
        want, rows, cls = po.torch_nms(rec[b], 0.1, iou, device=sem, return_index=True)
        ncand = int((rec[b][:, 4:] > np.float32(0.1)).sum())
        vanilla = 4 * ncand > (100000 if sem == "cuda" else 4000)
        assert_same_detections(outs[b].cpu().numpy(), want, ties_unordered=vanilla, what="%s/%s/%d" % (sem, profile, b))
        if not vanilla:
            assert np.array_equal(idxs[b].cpu().numpy(), rows * C + cls)


def test_torch_nms_matches_installed_torchvision_on_both_devices():
    tv = pytest.importorskip("torchvision")
    from pqdet_b200 import config, synth, tools
    C, size = 20, 512
    heads = synth.make_heads(3, C, size, "sparse", seed=21)
    pred = po.detect([h.numpy() for h in heads], C, synth.FPN_STRIDES)
    rec = torch.from_numpy(po.recover(pred, (size, size), np.array([float(size), float(size)], np.float32), "voc"))

    def reference_torch_nms(bb, thr, iou):                       # tools.py:540-566 with the live torchvision
        cs = bb[:, 4:]
        mask = cs > thr
        idx = mask.nonzero()
        keep = tv.ops.boxes.batched_nms(bb[:, :4][idx[:, 0]], cs[mask], idx[:, 1], iou)
        return torch.cat([bb[:, :4][idx[:, 0]][keep], cs[mask][keep, None], idx[:, 1][keep, None].float()], 1)

    for sem, dev in (("cpu", "cpu"), ("cuda", "cuda")):
        config.nms_semantics = sem
        for b in range(3):
            want = reference_torch_nms(rec[b].to(dev), 0.1, 0.45).cpu().numpy()
            got = tools.torch_nms(rec[b].cuda(), 0.1, 0.45).cpu().numpy()
            assert_same_detections(got, want, what="torchvision-%s/%d" % (dev, b))
            assert_same_detections(po.torch_nms(rec[b].numpy(), 0.1, 0.45, device=sem), want, what="oracle-%s" % dev)


def _fused_vs_chain(B, C, size, profile, kind, thr, iou, sem, seed, orig=None, strides=None, mode=None):
    from pqdet_b200 import config, fused, synth
    from pqdet_b200.interpreter import DetectionHead
    config.nms_semantics = sem
    hints = fused.StrategyHints()                                 # caller-held: every case starts on the fused kernel
    strides = strides or synth.FPN_STRIDES
    heads = synth.make_heads(B, C, size, profile, seed=seed, strides=strides)
    dheads = [h.cuda() for h in heads]
    if orig is None:
        orig = np.tile(np.array([[float(size), float(size)]], np.float32), (B, 1))
    opts = [dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in strides]
    decoded = DetectionHead(opts)(dheads).cpu().numpy()          # OUR decode: bit-exactness is defined on it
    want = eval_chain_oracle(None, strides, C, (size, size), orig, kind, thr, iou, sem,
                             mode=mode or "auto", decoded=decoded)
    kw = {}
    if mode:
        kw["nms_mode"] = mode
    dets = fused.decode_nms(dheads, strides, C, (size, size), cuda(orig), kind, thr, iou, return_index=True,
                            hints=hints, **kw)
    # the general path for every image must give the same rows (strategy='general', and what 'auto' switches to
    # after a mostly-overflowing call)
    # (small batches start on the large capacity class; both classes are asked for explicitly as well)
    for strat in ("general", "auto", "compact", "large"):
        alt = fused.decode_nms(dheads, strides, C, (size, size), cuda(orig), kind, thr, iou, return_index=True,
                               strategy=strat, hints=hints, **kw)
        if strat == "auto" and len(dets._spill) * 2 > B:
            assert alt.general_first and alt.images_via_general_path == B      # the hint routed the batch
        if strat in ("compact", "large"):
            assert alt.capacity == strat
        for b in range(B):
            assert torch.equal(alt[b], dets[b]) and torch.equal(alt.indices(b), dets.indices(b)), (strat, b)
    # decoded boxes themselves: within 1e-5 of the oracle's decode
    ref_dec = po.detect([h.numpy() for h in heads], C, strides)
    assert rel_close(decoded[..., :4], ref_dec[..., :4], 1e-5, scale=float(size))
    assert rel_close(decoded[..., 4:], ref_dec[..., 4:], 1e-5, scale=1e-30)
    total_cand = 0
    for b in range(B):
        w, rows, cls = want[b]
        rec_b = None
        got = dets[b].cpu().numpy()
        ncand = int(dets.host_meta()[1, b])
        total_cand += ncand
        limit = 100000 if sem == "cuda" else 4000
        vanilla = (mode == "vanilla") or (mode is None and 4 * ncand > limit)
        assert_same_detections(got, w, ties_unordered=vanilla, what="fused %s/%s img %d" % (profile, sem, b))
        if not vanilla:
            assert np.array_equal(dets.indices(b).cpu().numpy(), rows * C + cls)
    return dets, total_cand


@pytest.mark.parametrize("sem", ["cuda", "cpu"])
def test_fused_sparse_voc(sem):
    orig = np.array([[375., 500.], [333., 500.], [512., 512.], [500., 281.], [480., 640.], [512., 512.]], np.float32)
    dets, total = _fused_vs_chain(6, 20, 512, "sparse", "voc", 0.1, 0.45, sem, seed=31, orig=orig)
    assert total > 1500
    assert int(dets.host_meta()[2].sum()) == 0


def test_fused_predict_threshold_and_coco_iou():
    _fused_vs_chain(3, 20, 512, "sparse", "voc", 0.25, 0.45, "cpu", seed=32)        # predict.py:70
    _fused_vs_chain(2, 80, 608, "coco", "coco", 0.1, 0.65, "cuda", seed=33)         # yamls/coco.yaml:50, 19x19 level
    _fused_vs_chain(2, 20, 512, "sparse", "voc", 0.1, 0.45, "cuda", seed=34, strides=(8, 16, 32))   # PAN order


@pytest.mark.parametrize("sem", ["cuda", "cpu"])
def test_fused_dense_visdrone_overflows_to_general_path(sem):
    orig = np.array([[480., 480.], [360., 640.]], np.float32)
    dets, total = _fused_vs_chain(2, 10, 608, "dense", "visdrone", 0.1, 0.45, sem, seed=35, orig=orig)
    assert total > 2 * 2048                                     # beyond the on-chip candidate list
    assert len(dets._spill) == 2


def test_fused_forced_modes():
    _fused_vs_chain(2, 20, 512, "sparse", "voc", 0.1, 0.45, "cuda", seed=36, mode="vanilla")
    _fused_vs_chain(2, 20, 512, "sparse", "voc", 0.1, 0.45, "cpu", seed=37, mode="trick")


def test_fused_gauss_worst_case_small():
    # every row is a hit: the overflow path with ~10^5 candidates per image
    _fused_vs_chain(1, 20, 256, "gauss", "voc", 0.1, 0.45, "cuda", seed=38)


def test_fused_edge_cases():
    from pqdet_b200 import fused, synth
    C, size = 20, 512
    heads = [h.cuda() for h in synth.make_heads(2, C, size, "sparse", seed=39)]
    orig = torch.tensor([float(size), float(size)]).cuda()
    d = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.999999, 0.45)   # nothing passes
    assert d.counts.tolist() == [0, 0] and tuple(d[0].shape) == (0, 6)
    assert [tuple(t.shape) for t in d.to_reference_list()] == [(0,), (0,)]
    d = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 1.5, 0.45)
    assert d.counts.tolist() == [0, 0]
    d1 = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45)
    d2 = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45)
    for b in range(2):                                           # run-to-run determinism, bitwise
        assert torch.equal(d1[b], d2[b])


def test_fused_full_size_properties():
    """BASELINE config E shape (1024 x VOC-512): size-independent properties instead of the oracle:
    shard invariance (any split of the batch gives the same per-image rows -- the multi-GPU equality),
    sortedness, idempotence, and an oracle spot check on a few images."""
    from pqdet_b200 import fused, synth
    C, size, B = 20, 512, 1024
    heads = [h for h in synth.make_heads(B, C, size, "sparse", seed=0, device="cuda")]
    orig = torch.tensor([float(size), float(size)]).cuda()
    full = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45)
    assert int(full.host_meta()[2].sum()) == 0
    half = [fused.decode_nms([h[i:i + 512] for h in heads], synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45)
            for i in (0, 512)]
    for b in list(range(0, 1024, 37)) + [511, 512, 1023]:
        a = full[b]
        assert torch.equal(a, half[b // 512][b % 512])
        s = a[:, 4]
        assert bool((s[:-1] >= s[1:]).all()) and bool((s > 0.1).all())
        # idempotence: no two kept boxes of one class overlap above the threshold (NMS of the output keeps all of it)
        bx, cl = a[:, :4], a[:, 5]
        lt = torch.maximum(bx[:, None, :2], bx[None, :, :2])
        rb = torch.minimum(bx[:, None, 2:], bx[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        area = (bx[:, 2] - bx[:, 0]) * (bx[:, 3] - bx[:, 1])
        iou = inter / (area[:, None] + area[None, :] - inter)
        same = (cl[:, None] == cl[None, :]) & ~torch.eye(len(a), dtype=torch.bool, device=a.device)
        assert float(torch.where(same, iou, torch.zeros_like(iou)).max()) <= 0.45 + 1e-6
    counts = full.counts.to(torch.float32)
    assert 60 < float(counts.mean()) < 200
    from pqdet_b200.interpreter import DetectionHead
    opts = [dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in synth.FPN_STRIDES]
    for b in (0, 513, 1023):
        dec = DetectionHead(opts)([h[b:b + 1] for h in heads]).cpu().numpy()
        want = eval_chain_oracle(None, synth.FPN_STRIDES, C, (size, size), np.array([size, size], np.float32),
                                 "voc", 0.1, 0.45, "cuda", decoded=dec)[0][0]
        assert_same_detections(full[b].cpu().numpy(), want, what="full-size img %d" % b)


def test_fused_back_to_back_launches_share_output_buffers():
    """Consecutive launches on one stream overlap (programmatic dependent launch: a launch may start while the previous
    one drains) yet write the SAME output buffers: every launch must leave exactly its own results, whatever the
    order in which the two grids' CTAs finish.  Alternating a light and a heavy input makes the earlier launch the
    slower one."""
    from pqdet_b200 import _ops, synth
    C, size, B = 20, 512, 592
    dev = torch.device("cuda")
    orig = torch.tensor([float(size), float(size)], device=dev)
    sets = []
    for seed, thr in ((1, 0.1), (2, 0.02)):                     # the second set keeps many more candidates: slower images
        hs = synth.make_heads(B, C, size, "sparse", seed=seed, device=dev)
        sets.append((hs,) + _ops.make_heads(hs, synth.FPN_STRIDES, C, (size, size), orig, "voc", thr, 0.45, "auto_cuda", "tv_cuda"))
    want = []
    for _, h, keep in sets:
        det, idx, meta = _ops.decode_nms_fused(h, keep, 2048, True)
        torch.cuda.synchronize()
        want.append((det.clone(), idx.clone(), meta.clone()))
    out = _ops.alloc_fused_outputs(B, 2048, True, dev)
    for order in ((1, 0), (0, 1), (1, 1, 0), (0, 0, 1), (1, 0, 1, 0)):
        for which in order:
            _ops.decode_nms_fused(sets[which][1], sets[which][2], 2048, True, out=out)
        torch.cuda.synchronize()
        last = order[-1]
        cnt = out[2][:B]
        assert torch.equal(out[2][:3 * B], want[last][2][:3 * B]), order
        for b in range(0, B, 7):
            k = int(cnt[b])
            assert torch.equal(out[0][b, :k], want[last][0][b, :k]) and torch.equal(out[1][b, :k], want[last][1][b, :k]), (order, b)


def test_fused_launches_in_a_cuda_graph():
    """The overlapping launches (programmatic dependent launch) captured in one CUDA graph and replayed: same results."""
    from pqdet_b200 import _ops, synth
    B, C, size = 256, 20, 512
    dev = torch.device("cuda")
    orig = torch.tensor([float(size), float(size)], device=dev)
    sets = []
    for seed in (1, 2):
        hs = synth.make_heads(B, C, size, "sparse", seed=seed, device=dev)
        sets.append((hs,) + _ops.make_heads(hs, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda"))
    det, _, meta = _ops.decode_nms_fused(sets[1][1], sets[1][2], 2048, False)
    torch.cuda.synchronize()
    want_det, want_meta = det.clone(), meta.clone()
    out, out2 = _ops.alloc_fused_outputs(B, 2048, False, dev), _ops.alloc_fused_outputs(B, 2048, False, dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                               # warm-up on the capture stream arms the scheduler words
        for _ in range(2):
            _ops.decode_nms_fused(sets[0][1], sets[0][2], 2048, False, out=out)
            _ops.decode_nms_fused(sets[1][1], sets[1][2], 2048, False, out=out2)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        _ops.decode_nms_fused(sets[0][1], sets[0][2], 2048, False, out=out)
        _ops.decode_nms_fused(sets[1][1], sets[1][2], 2048, False, out=out2)
        _ops.decode_nms_fused(sets[1][1], sets[1][2], 2048, False, out=out)      # overwrites the first launch's rows
    for o in (out, out2):
        o[0].zero_()
        o[2][:3 * B].zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for o in (out, out2):
        assert torch.equal(o[2][:3 * B], want_meta[:3 * B])
        for b in range(0, B, 5):
            k = int(want_meta[b])
            assert torch.equal(o[0][b, :k], want_det[b, :k]), b


@pytest.mark.parametrize("size,C", [(320, 20), (352, 1), (416, 20), (480, 3), (544, 80), (576, 20)])
def test_fused_multi_scale_input_sizes(size, C):
    """The reference trains/evaluates at 320..608 (config.py:67): grids such as 11x11, 13x13, 19x19 exercise the
    partial 128-cell groups and the unaligned (scalar) objectness loads of the scan."""
    orig = np.array([[float(size) * 0.75, float(size)], [float(size), float(size) * 0.6]], np.float32)
    _fused_vs_chain(2, C, size, "sparse", "voc", 0.1, 0.45, "cuda", seed=size + C, orig=orig)


def test_fused_rectangular_heads_and_single_level():
    """Non-square grids (H != W) and a single-level model: oracle on our own decode."""
    from pqdet_b200 import config, fused
    from pqdet_b200.parser import Decode
    config.nms_semantics = "cuda"
    g = torch.Generator().manual_seed(77)
    C, B = 4, 3
    shapes = [(12, 20, 32), (24, 40, 16)]                         # (H, W, stride): 384 x 640 input
    heads = []
    for h, w, s in shapes:
        r = torch.randn((B, 3 * (5 + C), h, w), generator=g)
        r[:, 4::(5 + C)] -= 2.5                                   # objectness logits: ~8% of rows pass
        heads.append(r)
    orig = np.array([[384., 640.], [300., 500.], [640., 384.]], np.float32)
    for sub in (heads, heads[:1]):
        strides = [s for _, _, s in shapes][:len(sub)]
        dheads = [h.cuda() for h in sub]
        dec = torch.cat([Decode(C, s)(h).reshape(B, -1, 5 + C) for h, s in zip(dheads, strides)], dim=1).cpu().numpy()
        want = eval_chain_oracle(None, strides, C, (384, 640), orig, "coco", 0.1, 0.45, "cuda", decoded=dec)
        dets = fused.decode_nms(dheads, strides, C, (384, 640), cuda(orig), "coco", 0.1, 0.45, return_index=True)
        for b in range(B):
            w, rows, cls = want[b]
            ncand = int(dets.host_meta()[1, b])
            assert_same_detections(dets[b].cpu().numpy(), w, ties_unordered=4 * ncand > 100000, what="rect %d" % b)
            assert np.array_equal(dets.indices(b).cpu().numpy(), rows * C + cls)


@pytest.mark.parametrize("profile,C,size,kind", [("sparse", 20, 512, "voc"), ("dense", 10, 608, "visdrone")])
def test_host_buffer_path_equals_device_path(profile, C, size, kind):
    """pqdet_decode_nms_host: pinned host heads in, pinned host detections out, identical rows to the
    device-resident call (sparse: all images fused; dense: every image overflows -> general path)."""
    from pqdet_b200 import fused, synth
    B = 6
    heads = synth.make_heads(B, C, size, profile, seed=5)
    orig = torch.tensor([[375., 500.], [333., 500.], [float(size)] * 2, [500., 281.], [480., 480.], [300., 400.]])
    dev = fused.decode_nms([h.cuda() for h in heads], synth.FPN_STRIDES, C, (size, size), orig.cuda(), kind, 0.1, 0.45,
                           return_index=True)
    pinned = [h.pin_memory() for h in heads]
    for capacity in ("large", "compact"):                      # "large" is the default of the host path
        host = fused.decode_nms_host(pinned, synth.FPN_STRIDES, C, (size, size), orig, kind, 0.1, 0.45, return_index=True,
                                     capacity=capacity)
        rows = host.to_numpy_list()
        for b in range(B):
            assert np.array_equal(rows[b], dev[b].cpu().numpy()), (profile, capacity, b)
            assert np.array_equal(host.idx[b, :int(host.counts[b])].numpy(), dev.indices(b).cpu().numpy().astype(np.int32))
        assert int(host.counts.sum()) > 0


def test_host_buffer_path_rejects_pageable_memory():
    from pqdet_b200 import _lib, fused, synth
    heads = synth.make_heads(1, 20, 512, "sparse", seed=1)
    with pytest.raises(_lib.PqdetError):
        fused.decode_nms_host(heads, synth.FPN_STRIDES, 20, (512, 512), (512., 512.), "voc")


@pytest.mark.parametrize("sem", ["cuda", "cpu"])
def test_tie_break_identical_scores_and_shuffled_positions(sem):
    """SURVEY 8a' 'NMS tie-break': equal scores are visited in candidate order (row, then class); clusters of
    overlapping boxes with identical scores keep the first row, wherever the cluster sits in the tensor."""
    from pqdet_b200 import config, tools
    config.nms_semantics = sem
    rng = np.random.default_rng(5)
    N, C = 300, 4
    bb = np.zeros((N, 4 + C), np.float32)
    bb[:, :4] = np.array([1000., 1000., 1001., 1001.], np.float32)                 # far-away filler boxes
    rows = rng.permutation(N)[:100]
    for j, r in enumerate(rows):                                                    # 2 clusters x 50 boxes
        base = np.array([10., 10., 60., 60.], np.float32) if j < 50 else np.array([200., 200., 280., 260.], np.float32)
        bb[r, :4] = base + rng.uniform(-2, 2, 4).astype(np.float32)
        bb[r, 4 + (j % 2)] = np.float32(0.7)                                        # identical scores, two classes
    got, idx = tools.batched_torch_nms(cuda(bb[None]), 0.1, 0.45, return_index=True)
    want, wr, wc = po.torch_nms(bb, 0.1, 0.45, device=sem, return_index=True)
    assert_same_detections(got[0].cpu().numpy(), want, what="ties/" + sem)
    assert np.array_equal(idx[0].cpu().numpy(), wr * C + wc)
    first = {(j < 50, j % 2): min(r for jj, r in enumerate(rows) if (jj < 50) == (j < 50) and jj % 2 == j % 2)
             for j in range(100)}
    assert sorted(int(i) // C for i in idx[0].cpu().numpy()) == sorted(first.values())


def test_empty_batch_empty_rows_and_truncation():
    from pqdet_b200 import _lib, _ops, fused, synth, tools
    C, size = 20, 512
    heads = [h.cuda() for h in synth.make_heads(3, C, size, "sparse", seed=40)]
    orig = torch.tensor([float(size), float(size)]).cuda()
    d0 = fused.decode_nms([h[:0] for h in heads], synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45)
    assert len(d0) == 0 and d0.to_numpy_list() == []
    assert tuple(tools.torch_nms(torch.zeros((0, 4 + C), device="cuda"), 0.1, 0.45).shape) == (0,)
    # max_det smaller than the kept count: rows are truncated, counts[] still holds the true K, status says so
    full = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45)
    h, keep = _ops.make_heads(heads, synth.FPN_STRIDES, C, (size, size), orig, "voc", 0.1, 0.45, "auto_cuda", "tv_cuda")
    det, _, meta = _ops.decode_nms_fused(h, keep, 16, False)
    m = meta[:9].view(3, 3).cpu()
    assert m[0].tolist() == full.counts.tolist() and all(int(s) & _lib.ST_DET_TRUNCATED for s in m[2])
    for b in range(3):
        assert torch.equal(det[b], full[b][:16])


@pytest.mark.parametrize("profile,C,size,kind", [("dense", 10, 608, "visdrone"), ("gauss", 20, 256, "voc")])
def test_single_pass_select_equals_two_pass(profile, C, size, kind, monkeypatch):
    """General path: candidates staged per image in one read of the heads + gen_bucketize_kernel (default) against
    round 1's count / plan / re-evaluate route (PQDET_GEN_SELECT=twopass): identical rows and indices, also when the
    first capacity guess is too small (the retry loop of fused._general) and through the tools.torch_nms drop-in."""
    from pqdet_b200 import base_sample, config, fused, synth, tools
    from pqdet_b200.interpreter import DetectionHead
    config.nms_semantics = "cuda"
    B = 3
    heads = [h.cuda() for h in synth.make_heads(B, C, size, profile, seed=77)]
    orig = torch.tensor([[480., 480.], [360., 640.], [float(size)] * 2]).cuda()
    kw = dict(return_index=True, strategy="general")
    one = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, kind, 0.1, 0.45, **kw)
    opts = [dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in synth.FPN_STRIDES]
    rec = base_sample.RECOVER_BBOXES_REGISTER[kind](DetectionHead(opts)(heads), (size, size), orig)
    drop1, didx1 = tools.batched_torch_nms(rec, 0.1, 0.45, return_index=True, strategy="general")
    monkeypatch.setenv("PQDET_GEN_SELECT", "twopass")
    two = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), orig, kind, 0.1, 0.45, **kw)
    drop2, didx2 = tools.batched_torch_nms(rec, 0.1, 0.45, return_index=True, strategy="general")
    monkeypatch.delenv("PQDET_GEN_SELECT")
    assert int(one.counts.sum()) > 100
    for b in range(B):
        assert torch.equal(one[b], two[b]) and torch.equal(one.indices(b), two.indices(b)), b
        assert torch.equal(drop1[b], drop2[b]) and torch.equal(didx1[b], didx2[b]), b
        assert torch.equal(drop1[b], one[b])
