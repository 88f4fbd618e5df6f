"""Randomised sweep of the training path and of the section-8f components on a GPU box (not collected by pytest):

    python tests/stress_gpu_train.py [n_cases] [seed]

Per case: label assignment vs the oracle (bit-exact labels and GT lists), sparse targets vs dense labels (bit-exact
losses and gradients), the multi-level loss vs the oracle's torch formulation (1e-5 relative on the four losses when no
cell sits on the ignore threshold), head conv + decode (decoded == Decode(raw), raw within the TF32 bound), the AP
accumulator vs the oracle and the letterbox vs the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gpu_util import cuda  # noqa: E402
from oracle import ap_oracle, augment_oracle, loss_ref  # noqa: E402
from oracle import pqdet_oracle as po  # noqa: E402
from pqdet_b200 import _ops, augment, synth  # noqa: E402
from pqdet_b200.evaluator import DetectionAccumulator  # noqa: E402
from pqdet_b200.interpreter import DetectionHead  # noqa: E402
from pqdet_b200.train_dataset import DEFAULT_ANCHORS, LabelAssigner  # noqa: E402

VIS = [(9, 13), (25, 17), (16, 31), (47, 29), (32, 51), (83, 48), (61, 91), (131, 99), (210, 189)]


def train_case(rng, i):
    C = int(rng.choice([1, 2, 5, 10, 20, 80]))
    size = int(rng.choice([256, 288, 320, 416]))
    B = int(rng.integers(1, 5))
    lo, hi = (0, 6) if rng.random() < 0.5 else (10, 80)
    kind = str(rng.choice(["l1", "iou", "giou", "diou"]))
    anchors = DEFAULT_ANCHORS if rng.random() < 0.6 else VIS
    strides = (32, 16, 8) if rng.random() < 0.7 else (8, 16, 32)
    seed = int(rng.integers(0, 1 << 30))
    gts = synth.make_gt(B, C, size, max(lo, 0), hi, seed=seed)
    out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
    la = LabelAssigner(C, anchors=anchors)
    dense = la.create_label_batch(gts, out_sizes)
    sparse = la.create_sparse_batch(gts, out_sizes)
    wl, wg = po.create_label_batch(gts, out_sizes, C, anchors)
    tag = "case %d (C=%d size=%d B=%d GT %d-%d %s %s)" % (i, C, size, B, lo, hi, kind, strides)
    for s in range(3):
        assert np.array_equal(dense[s].cpu().numpy(), wl[s]), tag + " labels"
        assert np.array_equal(dense[3 + s].cpu().numpy(), wg[s]), tag + " gt lists"
    heads = synth.make_train_heads(B, C, size, seed=seed, strides=strides)
    head = DetectionHead([dict(classes=C, stride=s, bbox_loss=kind, ignore_thresh=0.5, l1_loss_gain=0.05) for s in strides])
    r1 = [h.cuda().requires_grad_(True) for h in heads]
    r2 = [h.cuda().requires_grad_(True) for h in heads]
    o1, o2 = head(r1, dense), head(r2, sparse)
    o1["loss"].mean().backward()
    o2["loss"].mean().backward()
    assert all(torch.equal(o1[k], o2[k]) for k in ("loss", "giou_loss", "conf_loss", "class_loss")), tag + " sparse loss"
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(r1, r2)), tag + " sparse grad"
    d3, g3 = head.loss_and_grad([h.cuda() for h in heads], dense)
    assert torch.equal(d3["loss"], o1["loss"].detach()) and all(torch.equal(a, b.grad) for a, b in zip(g3, r1)), tag
    total = np.zeros(4)
    ambiguous = False
    idx = {8: 0, 16: 1, 32: 2}
    for h, s in zip(heads, strides):
        lab_t, gt_t = torch.from_numpy(wl[idx[s]]), torch.from_numpy(wg[idx[s]])
        want, _ = loss_ref.yolo_layer_loss(h, lab_t, gt_t, C, s, kind, 0.5, 0.05, want_grad=False)
        total += np.array([float(w) for w in want])
        pred = loss_ref.decode_t(h, C, s)
        mx = loss_ref.iou_t(pred[..., None, 0:4], gt_t[:, None, None, None, :, :]).max(dim=-1)[0]
        ambiguous |= bool(((mx - 0.5).abs() < 1e-4).any())
    got = np.array([float(o1[k]) for k in ("loss", "giou_loss", "conf_loss", "class_loss")])
    tol = 1e-3 if ambiguous else 1e-5
    assert np.all(np.abs(got - total) <= tol * np.maximum(np.abs(total), 1e-30)), (tag, got, total)


def misc_case(rng, i):
    # head conv with awkward shapes
    Cin, H, W, C = int(rng.integers(5, 200)), int(rng.integers(3, 40)), int(rng.integers(3, 40)), int(rng.choice([1, 3, 20]))
    ACH = 3 * (5 + C)
    x = torch.randn((2, Cin, H, W), device="cuda")
    w = torch.randn((ACH, Cin), device="cuda") * 0.1
    b = torch.randn((ACH,), device="cuda") * 0.1
    dec, raw = _ops.head_conv_decode(x, w, b, C, 16.0, want_raw=True)
    ref = torch.einsum("bchw,oc->bohw", x.double(), w.double()) + b.double().view(1, -1, 1, 1)
    bound = 2.0 ** -9 * torch.einsum("bchw,oc->bohw", x.double().abs(), w.double().abs()) + 1e-5
    assert bool(((raw.double() - ref).abs() <= bound).all()), ("head conv", Cin, H, W, C)
    assert torch.equal(dec, _ops.decode_fwd(raw, C, 16.0)), ("head conv decode", Cin, H, W, C)
    # AP accumulator
    Cc = int(rng.choice([1, 4, 20]))
    dt = [np.float32, np.float64][int(rng.integers(0, 2))]
    data = synth.make_eval_set(int(rng.integers(5, 60)), Cc, 512, seed=int(rng.integers(0, 1 << 30)), gt_dtype=dt,
                               n_obj=(1, int(rng.integers(2, 60))))
    acc, orc = DetectionAccumulator(["c%d" % k for k in range(Cc)]), ap_oracle.ApOracle(Cc)
    for f, gt, diffs, dets in data:
        if len(dets) == 0:
            continue
        acc.add_detections(f, dets); acc.add_labels(f, gt, diffs)
        orc.add_detections(f, dets); orc.add_labels(f, gt, diffs)
    if acc.detections_count:
        with np.errstate(all="ignore"):
            want = orc.AP()
        assert np.array_equal(acc.AP().raw, want, equal_nan=True), ("AP", Cc, dt)
    # letterbox
    T = (int(rng.choice([64, 96, 160])), int(rng.choice([64, 128, 224])))
    imgs = [rng.integers(0, 256, (int(rng.integers(3, 300)), int(rng.integers(3, 300)), 3), dtype=np.uint8) for _ in range(3)]
    out, u8, _ = augment.letterbox_normalize(imgs, T, want_uint8=True)
    for k, im in enumerate(imgs):
        padded, _ = augment_oracle.resize_letterbox(im, T)
        assert np.array_equal(u8[k].cpu().numpy(), padded), ("letterbox", im.shape, T)
        assert np.array_equal(out[k].cpu().numpy(), augment_oracle.normalize_to_chw(padded, augment.VOC_MEAN, augment.VOC_STD))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    for i in range(n):
        train_case(rng, i)
        misc_case(rng, i)
    print("stress ok: %d training + %d misc cases" % (n, n))


if __name__ == "__main__":
    main()
