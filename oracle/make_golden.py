"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) on small seeded inputs.

Run in the build container (the reference cannot travel to the GPU box):

    python oracle/make_golden.py

The fixtures are committed; tests/test_oracle_golden.py pins oracle/ against them and the
`-m gpu` tests check the CUDA path against the same files.  Versions used are recorded in
tests/golden/MANIFEST.json.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from pqdet_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def gen_decode(ref):
    g = torch.Generator().manual_seed(11)
    d = {}
    C = 3
    shapes = {32: (2, 3), 16: (4, 5), 8: (7, 9)}                  # non-square: pins x/y orientation
    outs = []
    for s, (h, w) in shapes.items():
        raw = torch.randn((2, 3 * (5 + C), h, w), generator=g) * 1.5
        out = ref.Decode(C, s)(raw)
        d["raw_s%d" % s] = raw.numpy()
        d["out_s%d" % s] = out.numpy()
        outs.append(out.reshape(2, -1, 5 + C))
    d["concat"] = torch.cat(outs, dim=1).numpy()                  # DetectionModel eval concat order
    d["zeros_s8_cell_1_2"] = ref.Decode(20, 8)(torch.zeros(1, 75, 2, 3))[0, 1, 2, 0].numpy()
    d["num_classes"] = np.int64(C)
    return d


def gen_recover(ref):
    g = torch.Generator().manual_seed(12)
    C, B, N = 4, 3, 64
    pred = torch.empty((B, N, 5 + C))
    c = torch.rand((B, N, 2), generator=g) * 420 + 40
    wh = torch.rand((B, N, 2), generator=g) * 300 - 20
    pred[..., 0:2] = c - wh / 2
    pred[..., 2:4] = c + wh / 2
    pred[..., 4:] = torch.rand((B, N, 1 + C), generator=g)
    inp = torch.tensor([512.0, 512.0])
    orig = torch.tensor([[375.0, 500.0], [500.0, 333.0], [281.0, 500.0]])
    d = {"pred": pred.numpy(), "input_size": inp.numpy(), "orig": orig.numpy()}
    for kind in ("voc", "coco", "visdrone"):
        d["out_" + kind] = ref.RECOVER[kind](pred.clone(), inp, orig).numpy()
    d["out_voc_1d"] = ref.RECOVER["voc"](pred.clone(), inp, orig[0]).numpy()
    inp2 = torch.tensor([608.0, 416.0])
    d["input_size2"] = inp2.numpy()
    d["out_voc_rect"] = ref.RECOVER["voc"](pred.clone(), inp2, orig).numpy()
    return d


def _cluster_boxes(g, n_obj, per, C, size=500.0):
    ctr = torch.rand((n_obj, 2), generator=g) * (size - 100) + 50
    wh = torch.rand((n_obj, 2), generator=g) * 120 + 20
    rows = []
    for k in range(n_obj):
        jit = torch.randn((per, 4), generator=g) * 4
        b = torch.cat([ctr[k] - wh[k] / 2, ctr[k] + wh[k] / 2]).unsqueeze(0) + jit
        sc = torch.rand((per, C), generator=g) * 0.12
        cls = int(torch.randint(0, C, (1,), generator=g))
        sc[:, cls] = torch.rand((per,), generator=g) * 0.7 + 0.3
        rows.append(torch.cat([b, sc], dim=1))
    return torch.cat(rows, dim=0)


def gen_nms(ref):
    g = torch.Generator().manual_seed(13)
    d = {}
    cases = {
        "small": (_cluster_boxes(g, 6, 20, 5), 0.1, 0.45),
        "coco_iou": (_cluster_boxes(g, 10, 30, 8), 0.1, 0.65),
        "predict_thr": (_cluster_boxes(g, 5, 25, 3), 0.25, 0.45),
        "vanilla_cpu": (_cluster_boxes(g, 40, 40, 6), 0.05, 0.45),     # > 1000 candidates
        "empty": (_cluster_boxes(g, 3, 10, 4) * torch.tensor([1, 1, 1, 1, .01, .01, .01, .01]), 0.5, 0.45),
    }
    tie = _cluster_boxes(g, 2, 25, 3)
    tie[:, 4:] = 0.0
    tie[:25, 4] = 0.8
    tie[25:, 5] = 0.8                                            # all scores equal inside a cluster
    cases["ties"] = (tie, 0.1, 0.45)
    neg = _cluster_boxes(g, 4, 12, 3)
    neg[::5, 2] = neg[::5, 0] - 3.0                              # degenerate (negative width) boxes
    cases["degenerate"] = (neg, 0.1, 0.45)
    for name, (bb, thr, iou) in cases.items():
        out = ref.tools.torch_nms(bb, thr, iou)
        d[name + "_in"] = bb.numpy()
        d[name + "_thr"] = np.float64(thr)
        d[name + "_iou"] = np.float64(iou)
        d[name + "_out"] = out.numpy()
        d[name + "_ncand"] = np.int64(int((bb[:, 4:] > thr).sum()))
    return d


def gen_label_and_loss(ref):
    d = {}
    C, size, B = 4, 128, 3
    gts = synth.make_gt(B, C, size, 1, 6, seed=21)
    gts[1][:, 5] = 0.35                                          # mixup weight column
    gts[2] = np.concatenate([gts[2], gts[2][:1] + np.array([1, 1, 1, 1, 0, 0], np.float32)])  # cell collision
    gts[2][-1, 4] = (gts[2][0, 4] + 1) % C
    ds = rh.make_label_dataset(C)
    out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
    per = [ds.create_label(gb, out_sizes) for gb in gts]
    batch = ref.collate_batch([(np.zeros((1,), np.float32),) + p for p in per])
    labels = [b.numpy() for b in batch[1:4]]
    gtl = [b.numpy() for b in batch[4:7]]
    d["num_classes"] = np.int64(C)
    d["size"] = np.int64(size)
    d["anchors"] = np.array(rh.VOC_ANCHORS, np.float32)
    d["gt_counts"] = np.array([len(x) for x in gts], np.int64)
    gpad = np.zeros((B, max(len(x) for x in gts), 6), np.float32)
    for b, x in enumerate(gts):
        gpad[b, :len(x)] = x
    d["gt"] = gpad
    for i, s in enumerate((8, 16, 32)):
        d["label_s%d" % s] = labels[i]
        d["gtlist_s%d" % s] = gtl[i]
    # an image with no GT at all (labels stay background, lists pad to one zero row)
    per0 = ds.create_label(np.zeros((0, 6), np.float32), out_sizes)
    d["empty_label_s8"] = per0[0]
    heads = synth.make_train_heads(B, C, size, seed=22, strides=(8, 16, 32))
    for i, s in enumerate((8, 16, 32)):
        d["raw_s%d" % s] = heads[i].numpy()
        for kind in ("l1", "iou", "giou", "diou"):
            opt = dict(classes=C, stride=s, bbox_loss=kind, ignore_thresh=0.5, l1_loss_gain=0.05)
            raw = heads[i].clone().requires_grad_(True)
            out = ref.YOLOLayer(opt)(raw, (t(labels[i]), t(gtl[i])))
            out[0].sum().backward()
            d["loss_%s_s%d" % (kind, s)] = np.array([float(o) for o in out], np.float32)
            d["grad_%s_s%d" % (kind, s)] = raw.grad.numpy()
        # standalone loss_per_scale on a decoded tensor: gradient w.r.t. pred
        pred = ref.Decode(C, s)(heads[i]).detach().requires_grad_(True)
        opt = dict(classes=C, stride=s, bbox_loss="giou", ignore_thresh=0.5, l1_loss_gain=0.05)
        out = ref.loss_per_scale(pred, t(labels[i]), t(gtl[i]), opt)
        out[0].sum().backward()
        d["predgrad_giou_s%d" % s] = pred.grad.numpy()
    # ciou always raises (SURVEY.md box, item 6)
    try:
        opt = dict(classes=C, stride=8, bbox_loss="ciou", ignore_thresh=0.5, l1_loss_gain=0.05)
        ref.YOLOLayer(opt)(heads[0], (t(labels[0]), t(gtl[0])))
        d["ciou_raises"] = np.int64(0)
    except RuntimeError as e:
        d["ciou_raises"] = np.int64(1 if "NaN in loss" in str(e) else 0)
    return d


def gen_iou(ref):
    g = torch.Generator().manual_seed(14)
    c = torch.rand((200, 2, 2), generator=g) * 100
    wh = torch.rand((200, 2, 2), generator=g) * 60 + 1
    b = torch.cat([c - wh / 2, c + wh / 2], dim=-1)
    b1, b2 = b[:, 0], b[:, 1]
    d = {"b1": b1.numpy(), "b2": b2.numpy()}
    for name in ("iou_calc3", "giou", "diou", "ciou"):
        d[name] = getattr(ref.tools, name)(b1, b2).numpy()
    d["iou_calc3_bcast"] = ref.tools.iou_calc3(b1[:7, None, :], b2[None, :9, :]).numpy()
    d["iou_calc1"] = ref.tools.iou_calc1(b1.numpy(), b2.numpy())
    xywh1 = np.array([50.0, 40.0, 30.0, 20.0], np.float32)
    xywh2 = np.array([[52, 44, 10, 13], [52, 44, 33, 23], [60, 60, 116, 90]], np.float64)
    d["xywh1"], d["xywh2"] = xywh1, xywh2
    d["iou_xywh_numpy"] = ref.tools.iou_xywh_numpy(xywh1, xywh2)
    return d


def gen_ap(ref):
    """Evaluator.add_detections / add_labels / AP on synthetic evaluation sets (float32 and float64 GT)."""
    import contextlib
    import io
    d = {}
    for tag, dt, C, n, seed in (("f32", np.float32, 5, 40, 3), ("f64", np.float64, 20, 60, 4)):
        ev = rh.make_evaluator(["c%d" % i for i in range(C)])
        data = synth.make_eval_set(n, C, 512, seed=seed, gt_dtype=dt)
        for f, gt, diffs, dets in data:
            ev.add_detections(f, dets)
            ev.add_labels(f, gt, diffs)
        with contextlib.redirect_stderr(io.StringIO()):
            ap = ev.AP()
        d["raw_" + tag] = ap.raw
        d["mAPs_" + tag] = np.asarray(ap.mAPs)
        d["APs_" + tag] = np.asarray(ap.APs)
        d["AP_" + tag] = np.float64(ap.AP)
        d["cfg_" + tag] = np.array([n, C, 512, seed], np.int64)
    return d


def gen_letterbox(ref):
    """augment.Resize -> Normalize -> HWCtoCHW (= ToTensor without the device copy) of the reference on small random
    uint8 images: up-scaling, down-scaling, tall, wide, already-square."""
    import sys as _sys
    _sys.path.insert(0, rh.REFERENCE_ROOT)
    from dataset import augment as raug
    rng = np.random.default_rng(21)
    d = {}
    shapes = [(37, 53), (120, 45), (64, 64), (23, 111), (150, 97), (9, 14)]
    for i, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bb = rng.uniform(0, min(h, w), (4, 4)).astype(np.float32)
        rimg, rbb = raug.Resize((64, 96))(img.copy(), bb.copy())
        nimg, _ = raug.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])(rimg, rbb)
        timg, _ = raug.HWCtoCHW()(nimg, rbb)
        d["img%d" % i], d["bb%d" % i] = img, bb
        d["resized%d" % i], d["rbb%d" % i], d["tensor%d" % i] = rimg, rbb, timg
    d["n"] = np.int64(len(shapes))
    d["target_hw"] = np.array([64, 96], np.int64)
    # eval_augment_visdrone: ResizeRatio(1.25) -> PadNearestDivisor -> Normalize -> HWCtoCHW
    rng = np.random.default_rng(22)
    vshapes = [(37, 53), (100, 80), (51, 26)]
    for i, (h, w) in enumerate(vshapes):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bb = rng.uniform(0, min(h, w), (3, 4)).astype(np.float32)
        r, rbb = raug.ResizeRatio(1.25)(img.copy(), bb.copy())
        r, rbb = raug.PadNearestDivisor()(r, rbb)
        nimg, _ = raug.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])(r, rbb)
        timg, _ = raug.HWCtoCHW()(nimg, rbb)
        d["vis_img%d" % i], d["vis_bb%d" % i] = img, bb
        d["vis_padded%d" % i], d["vis_rbb%d" % i], d["vis_tensor%d" % i] = r, rbb, timg
    d["vis_n"] = np.int64(len(vshapes))
    return d


def main():
    import torchvision
    ref = rh.load()
    os.makedirs(OUT, exist_ok=True)
    parts = {"decode": gen_decode, "recover": gen_recover, "nms": gen_nms,
             "train": gen_label_and_loss, "iou": gen_iou, "ap": gen_ap, "letterbox": gen_letterbox}
    sizes = {}
    for name, fn in parts.items():
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **fn(ref))
        sizes[name] = os.path.getsize(path)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "reference": "eleflea/PQDet @ /root/reference (unmodified)",
                   "torch": torch.__version__, "torchvision": torchvision.__version__,
                   "numpy": np.__version__, "device": "cpu", "bytes": sizes}, f, indent=1)
    print(sizes)


if __name__ == "__main__":
    main()
