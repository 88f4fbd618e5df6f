"""Pins oracle/ against the LIVE reference (imported from /root/reference through
oracle/ref_harness.py) on seeded random inputs.  Skipped where the reference is absent (GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from oracle import pqdet_oracle as po
from oracle import loss_ref
from pqdet_b200 import synth
from conftest import rel_close

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return rh.load()


def test_eval_chain_matches_reference(ref):
    C, size, B = 20, 512, 3
    heads = synth.make_heads(B, C, size, "sparse", seed=5)
    outs = [ref.Decode(C, s)(h) for h, s in zip(heads, synth.FPN_STRIDES)]
    pred = torch.cat([o.reshape(B, -1, 5 + C) for o in outs], dim=1)
    mine = po.detect([h.numpy() for h in heads], C, synth.FPN_STRIDES)
    assert mine.shape == (B, 16128, 25)                      # export/onnx_exporter.py:373 known answer
    assert rel_close(mine[..., :4], pred[..., :4].numpy(), 1e-5, scale=float(size))
    assert rel_close(mine[..., 4:], pred[..., 4:].numpy(), 1e-5, scale=1e-30)
    inp = torch.tensor([512.0, 512.0])
    orig = torch.tensor([[375.0, 500.0], [333.0, 500.0], [512.0, 512.0]])
    for kind in ("voc", "coco", "visdrone"):
        want = ref.RECOVER[kind](pred.clone(), inp, orig).numpy()
        got = po.recover(pred.numpy(), inp.numpy(), orig.numpy(), kind)
        assert np.array_equal(got, want), kind
    rec = ref.RECOVER["voc"](pred.clone(), inp, orig)
    for b in range(B):
        want = ref.tools.torch_nms(rec[b], 0.1, 0.45).numpy()
        got = po.torch_nms(rec[b].numpy(), 0.1, 0.45, device="cpu")
        assert np.array_equal(got, want)


def test_head_conv_oracle_matches_the_reference_model_head():
    """the oracle's head convolution + decode against the reference's own modules - the nn.Conv2d the cfg
    parser builds in front of a [yolo] layer and its Decode - on CPU (fp32 accumulation)."""
    import os
    ref = rh.load()
    from model.interpreter import DetectionModel
    m = DetectionModel(os.path.join(rh.REFERENCE_ROOT, "model", "cfg", "regnetx-600m-fpn.cfg")).eval()
    i = [k for k, l in enumerate(m.module_list) if l._type == "yolo"][0]
    block, yolo = m.module_list[i - 1], m.module_list[i]
    conv = block.conv
    assert len(block) == 1 and conv.kernel_size == (1, 1) and conv.out_channels == 75      # the plain linear head conv
    torch.manual_seed(3)
    x = torch.randn(2, conv.in_channels, 4, 6)
    with torch.no_grad():
        raw_ref = block(x)
        dec_ref = yolo(raw_ref)
    raw = po.head_conv(x.numpy(), conv.weight.detach().numpy(), conv.bias.detach().numpy())
    assert np.all(np.abs(raw - raw_ref.numpy()) <= po.head_conv_error_bound(x.numpy(), conv.weight.detach().numpy(), 2.0 ** -22))
    dec = po.decode(raw.astype(np.float32), 20, yolo.opt["stride"] if hasattr(yolo, "opt") else 32)
    assert rel_close(dec[..., :4], dec_ref.numpy()[..., :4], 1e-4, scale=float(32 * 6))
    assert rel_close(dec[..., 4:], dec_ref.numpy()[..., 4:], 1e-4, scale=1e-30)


def test_dense_nms_takes_vanilla_path_on_cpu(ref):
    C, size = 10, 608
    heads = synth.make_heads(1, C, size, "dense", seed=2)
    pred = torch.from_numpy(po.detect([h.numpy() for h in heads], C, synth.FPN_STRIDES))
    rec = ref.RECOVER["visdrone"](pred.clone(), torch.tensor([608.0, 608.0]), torch.tensor([[480.0, 480.0]]))
    want = ref.tools.torch_nms(rec[0], 0.1, 0.45).numpy()
    got = po.torch_nms(rec[0].numpy(), 0.1, 0.45, device="cpu")
    assert int((rec[0][:, 4:] > 0.1).sum()) > 1000
    assert got.shape == want.shape
    # vanilla re-sorts kept boxes with an unstable sort: compare as sets of rows + score order
    assert np.array_equal(got[:, 4], want[:, 4])
    assert set(map(bytes, got)) == set(map(bytes, want))


@pytest.mark.parametrize("kind", ["l1", "iou", "giou", "diou"])
def test_loss_matches_reference(ref, kind):
    C, size, B = 20, 256, 2
    gts = synth.make_gt(B, C, size, 1, 12, seed=3)
    out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
    labels, gtl = po.create_label_batch(gts, out_sizes, C, rh.VOC_ANCHORS)
    ds = rh.make_label_dataset(C)
    per = [ds.create_label(g, out_sizes) for g in gts]
    batch = ref.collate_batch([(np.zeros((1,), np.float32),) + p for p in per])
    for i in range(3):
        assert np.array_equal(labels[i], batch[1 + i].numpy())
        assert np.array_equal(gtl[i], batch[4 + i].numpy())
    heads = synth.make_train_heads(B, C, size, seed=5, strides=(8, 16, 32))
    for li, s in enumerate((8, 16, 32)):
        opt = dict(classes=C, stride=s, bbox_loss=kind, ignore_thresh=0.5, l1_loss_gain=0.05)
        raw = heads[li].clone().requires_grad_(True)
        out = ref.YOLOLayer(opt)(raw, (batch[1 + li], batch[4 + li]))
        out[0].sum().backward()
        mine, grad = loss_ref.yolo_layer_loss(heads[li], torch.from_numpy(labels[li]),
                                              torch.from_numpy(gtl[li]), C, s, kind, 0.5, 0.05)
        for a, b in zip(mine, out):
            assert rel_close(a.numpy(), b.detach().numpy(), 1e-6, scale=1e-30)
        assert rel_close(grad.numpy(), raw.grad.numpy(), 1e-6, scale=float(raw.grad.abs().max()))


def test_nms_oracle_matches_installed_torchvision_cpu():
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(0)
    for trial in range(6):
        n = 300
        c = torch.rand((n, 2), generator=g) * 80
        wh = torch.rand((n, 2), generator=g) * 40 + 2
        boxes = torch.cat([c - wh / 2, c + wh / 2], dim=1)
        scores = torch.rand((n,), generator=g)
        cls = torch.randint(0, 3, (n,), generator=g)
        for thr in (0.45, 0.65, 0.3):
            want = tv.ops.nms(boxes, scores, thr).numpy()
            got = po.nms_plain(boxes.numpy(), scores.numpy(), thr, round_mode=0)
            assert np.array_equal(got, want)
            want_b = tv.ops.batched_nms(boxes, scores, cls, thr).numpy()
            got_b = po.batched_nms(boxes.numpy(), scores.numpy(), cls.numpy(), thr, device="cpu")
            assert np.array_equal(got_b, want_b)


def test_ap_oracle_vs_live_reference_evaluator():
    import contextlib
    import io
    from oracle import ap_oracle
    from pqdet_b200 import synth
    for dt, C, n, seed in ((np.float32, 7, 30, 11), (np.float64, 3, 25, 12), (np.int64, 4, 20, 13)):
        ev = rh.make_evaluator(["c%d" % i for i in range(C)])
        orc = ap_oracle.ApOracle(C)
        for f, gt, diffs, dets in synth.make_eval_set(n, C, 512, seed=seed, gt_dtype=dt, n_obj=(1, 25)):
            ev.add_detections(f, dets); ev.add_labels(f, gt, diffs)
            orc.add_detections(f, dets); orc.add_labels(f, gt, diffs)
        with contextlib.redirect_stderr(io.StringIO()):
            ap = ev.AP()
        assert np.array_equal(ap.raw, orc.AP()), dt


def test_resize_oracle_is_bit_exact_against_cv2():
    """The third-party arithmetic on this path: cv2.resize(INTER_LINEAR) on uint8 (OpenCV fixed-point bilinear)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import augment_oracle as ao
    rng = np.random.default_rng(3)
    for _ in range(40):
        sh, sw = int(rng.integers(5, 400)), int(rng.integers(5, 400))
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        T = int(rng.choice([64, 160, 320, 512]))
        r = min(T / sw, T / sh)
        dw, dh = max(round(r * sw), 1), max(round(r * sh), 1)
        assert np.array_equal(ao.resize_linear_u8(img, dw, dh), cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("kind,C,size,dist", [("voc", 20, 512, "sparse"), ("coco", 80, 320, "sparse"),
                                              ("visdrone", 10, 608, "dense")])
def test_cpu_path_port_is_the_reference_sequence(ref, kind, C, size, dist):
    """oracle/cpu_path.py is the denominator of every speed-up bench.py reports (`--impl reference`, cpu_baseline).
    Pin it bit for bit to the live reference's predict.py:33-45 sequence: Decode per level -> cat
    (model/interpreter.py:72-75) -> RECOVER_BBOXES_REGISTER[kind] -> tools.torch_nms per image, for all three
    affines, sequential and image-parallel (threads and processes)."""
    from oracle import cpu_path
    B = 3 if dist == "sparse" else 2
    heads = synth.make_heads(B, C, size, dist, seed=17)
    strides = synth.FPN_STRIDES
    inp = torch.tensor([float(size), float(size)])
    orig = torch.tensor([[375.0, 500.0], [333.0, 500.0], [float(size), float(size)]])[:B]
    with torch.no_grad():
        outs = [ref.Decode(C, s)(h) for h, s in zip(heads, strides)]
        pred = torch.cat([o.reshape(B, -1, 5 + C) for o in outs], dim=1)
        # the port's decode / recover stages on their own, bit for bit
        mine = torch.cat([cpu_path.decode_t(h, C, s).reshape(B, -1, 5 + C) for h, s in zip(heads, strides)], dim=1)
        assert torch.equal(mine, pred)
        rec = ref.RECOVER[kind](pred.clone(), inp, orig)
        assert torch.equal(cpu_path.recover_t(pred, inp, orig, kind), rec)
        want = [ref.tools.torch_nms(rec[b], 0.1, 0.45) for b in range(B)]
    assert sum(w.shape[0] for w in want) > 0
    runs = {
        "sequential": cpu_path.eval_chain(heads, strides, C, inp, orig, kind, 0.1, 0.45),
        "threads": cpu_path.eval_chain_image_parallel(heads, strides, C, inp, orig, kind, 0.1, 0.45, workers=2),
    }
    if hasattr(cpu_path, "eval_chain_process_parallel"):
        runs["processes"] = cpu_path.eval_chain_process_parallel(heads, strides, C, inp, orig, kind, 0.1, 0.45,
                                                                 workers=2)
    for name, got in runs.items():
        assert len(got) == B, name
        for b in range(B):
            assert got[b].shape == want[b].shape, (name, b)
            assert torch.equal(got[b], want[b]), (name, b)


def test_numpy_nms_and_ciou_oracles_match_the_reference(ref):
    """Row a13 / a8: the oracle's restatements of tools.nms (hard and soft, tools.py:507-538), tools.iou_calc1 and
    tools.ciou (value and autograd gradient) against the live reference."""
    rng = np.random.default_rng(8)
    n = 120
    c = rng.random((n, 2)) * 100
    wh = rng.random((n, 2)) * 40 + 4
    bb = np.concatenate([c - wh / 2, c + wh / 2, rng.random((n, 1)), rng.integers(0, 4, (n, 1))], axis=1).astype(np.float32)
    assert np.array_equal(po.iou_calc1(bb[:50, None, :4], bb[None, 50:90, :4]), ref.tools.iou_calc1(bb[:50, None, :4], bb[None, 50:90, :4]))
    for method, thr in (("nms", 0.3), ("soft-nms", 0.3), ("soft-nms", 0.05), ("nms", 0.95)):
        want = ref.tools.nms(bb.copy(), thr, 0.45, sigma=0.3, method=method)
        got = po.numpy_nms(bb.copy(), thr, 0.45, sigma=0.3, method=method)
        assert want.shape == got.shape and len(want) > 0
        # the reference walks the classes in set order: compare class by class
        for cls in np.unique(bb[:, 5]):
            assert np.array_equal(got[got[:, 5] == cls], want[want[:, 5] == cls]), (method, cls)
    g = torch.Generator().manual_seed(4)
    p = (torch.rand((64, 2), generator=g) * 50)
    p = torch.cat([p, p + torch.rand((64, 2), generator=g) * 30 + 1], dim=1).requires_grad_(True)
    q = (torch.rand((64, 2), generator=g) * 50)
    q = torch.cat([q, q + torch.rand((64, 2), generator=g) * 30 + 1], dim=1).requires_grad_(True)
    a = ref.tools.ciou(p, q)
    a.sum().backward()
    p2, q2 = p.detach().clone().requires_grad_(True), q.detach().clone().requires_grad_(True)
    b = loss_ref.ciou_t(p2, q2)
    b.sum().backward()
    assert torch.equal(a, b)
    # same formula, differently shaped autograd graph: the accumulation order of the gradient terms differs
    assert torch.allclose(p.grad, p2.grad, rtol=1e-5, atol=1e-8) and torch.allclose(q.grad, q2.grad, rtol=1e-5, atol=1e-8)
