"""TEST INFRASTRUCTURE ONLY -- the checker, never the product path.

CPU restatement (numpy fp32, one rounding per operation, no contraction) of PQDet's detection
hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; pqdet_b200/ never does (tests/test_no_oracle_in_product.py checks).

Pinned against: (1) tests/golden/*.npz -- outputs of the unmodified reference generated in the
build container by oracle/make_golden.py (the reference ships no tests or golden vectors of its
own); (2) the live reference where /root/reference exists
(tests/test_oracle_vs_reference.py).  The NMS arithmetic lives in a third-party dependency
(torchvision, requirements.txt:3) and is restated in oracle/nms_oracle.c; it is additionally pinned
against the installed torchvision on CPU (here) and CUDA (GPU box) by the tests.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

F = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force: bool = False) -> str:
    """Compile oracle/nms_oracle.c with gcc (oracle/Makefile)."""
    out = os.path.join(_HERE, "_build", "libpqdet_oracle.so")
    src = os.path.join(_HERE, "nms_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        i64p = ctypes.POINTER(ctypes.c_int64)
        f32p = ctypes.POINTER(ctypes.c_float)
        for name in ("pq_oracle_nms_ordered", "pq_oracle_nms_mask"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int64
            fn.argtypes = [f32p, i64p, ctypes.c_int64, ctypes.c_double, ctypes.c_int, i64p]
        lib.pq_oracle_trick_offsets.restype = None
        lib.pq_oracle_trick_offsets.argtypes = [f32p, i64p, ctypes.c_int64, f32p]
        lib.pq_oracle_ignore_mask.restype = None
        lib.pq_oracle_ignore_mask.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_int64,
                                              ctypes.c_float, ctypes.POINTER(ctypes.c_ubyte)]
        _LIB = lib
    return _LIB


def _f32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


# --------------------------------------------------------------------------------------------
# a1/a2  decode  (model/parser.py:185-192, 206-235)
# --------------------------------------------------------------------------------------------
def sigmoid(x):
    x = np.asarray(x, dtype=F)
    with np.errstate(over="ignore"):
        return (F(1.0) / (F(1.0) + np.exp(-x))).astype(F)


def decode(raw: np.ndarray, num_classes: int, stride: int) -> np.ndarray:
    """raw (B, A*(5+C), H, W) NCHW -> (B, H, W, A, 5+C).

    cell centre = index + 0.5 with x along W and y along H (parser.py:185-192, the names are
    swapped there on purpose); xymin = (centre - exp(r01))*stride, xymax = (centre + exp(r23))*stride
    (parser.py:226-228); conf/prob = sigmoid (parser.py:230-232).
    """
    raw = np.ascontiguousarray(raw, dtype=F)
    B, CH, H, W = raw.shape
    ch = 5 + num_classes
    A = CH // ch
    r = raw.reshape(B, A, ch, H, W).transpose(0, 3, 4, 1, 2)          # (B,H,W,A,ch) view
    gx = (np.arange(W, dtype=F) + F(0.5)).reshape(1, 1, W, 1)
    gy = (np.arange(H, dtype=F) + F(0.5)).reshape(1, H, 1, 1)
    s = F(stride)
    out = np.empty((B, H, W, A, ch), dtype=F)
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp(r[..., 0:4]).astype(F)
        out[..., 0] = (gx - e[..., 0]) * s
        out[..., 1] = (gy - e[..., 1]) * s
        out[..., 2] = (gx + e[..., 2]) * s
        out[..., 3] = (gy + e[..., 3]) * s
    out[..., 4:] = sigmoid(r[..., 4:])
    return out


def detect(raws, num_classes: int, strides) -> np.ndarray:
    """Eval branch of DetectionModel.forward (model/interpreter.py:75-76): per-level decode,
    flatten to (B, H*W*A, 5+C) and concatenate along rows in the order given."""
    outs = [decode(r, num_classes, s) for r, s in zip(raws, strides)]
    return np.concatenate([o.reshape(o.shape[0], -1, o.shape[-1]) for o in outs], axis=1)


def head_conv(x: np.ndarray, weight: np.ndarray, bias) -> np.ndarray:
    """The `filters = A*(5+C), size = 1, stride = 1, activation = linear` convolution in front of every [yolo] layer
    (model/cfg/regnetx-600m-fpn.cfg:646-651; built as nn.Conv2d by model/parser.py:385-401): a per-cell matrix product,
    accumulated in float64 (the reference's ATen kernel accumulates in fp32 on CPU and in TF32 products / fp32
    accumulation on CUDA; both are compared to this within their precision).  x (B,Cin,H,W), weight (O,Cin[,1,1]),
    bias (O,) or None -> raw (B,O,H,W) float64."""
    w = np.asarray(weight, dtype=np.float64).reshape(weight.shape[0], -1)
    raw = np.einsum("bchw,oc->bohw", np.asarray(x, dtype=np.float64), w, optimize=True)
    if bias is not None:
        raw = raw + np.asarray(bias, dtype=np.float64).reshape(1, -1, 1, 1)
    return raw


def head_conv_error_bound(x: np.ndarray, weight: np.ndarray, rel: float) -> np.ndarray:
    """|error| allowed for a head convolution whose operands carry a relative error `rel` each (2^-10 for TF32
    truncation, ~2^-24 for fp32 accumulation noise): 2 * rel * sum_c |x_c * w_c| + 1e-5."""
    w = np.abs(np.asarray(weight, dtype=np.float64)).reshape(weight.shape[0], -1)
    return 2.0 * rel * np.einsum("bchw,oc->bohw", np.abs(np.asarray(x, dtype=np.float64)), w, optimize=True) + 1e-5


# --------------------------------------------------------------------------------------------
# a5  recover  (dataset/base_sample.py:98-139, voc_sample.py:92-95, coco_sample.py:97-100,
#               visdrone_sample.py:84-88)
# --------------------------------------------------------------------------------------------
def affine_params(kind: str, input_size, original_size):
    """-> delta (B,2) in (h,w) order, ratio (B,1).  All fp32; round is half-to-even."""
    inp = np.asarray(input_size, dtype=F).reshape(1, 2)
    org = np.asarray(original_size, dtype=F).reshape(-1, 2)
    if kind in ("voc", "coco"):
        ratio = (inp / org).min(axis=-1, keepdims=True).astype(F)
        delta = ((inp - np.rint(ratio * org).astype(F)) / F(2.0)).astype(F)
        return np.floor(delta).astype(F), ratio
    if kind == "visdrone":
        r = F(1.25)
        inp2 = (np.ceil((r * org) / F(32.0)) * F(32.0)).astype(F)
        delta = ((inp2 - r * org) / F(2.0)).astype(F)
        return np.floor(delta).astype(F), np.full((org.shape[0], 1), r, dtype=F)
    raise ValueError(kind)


def recover(pred: np.ndarray, input_size, original_size, kind: str = "voc") -> np.ndarray:
    """(B,N,5+C) decoded -> (B,N,4+C): undo letterbox, clip, fold conf into class probs."""
    pred = np.asarray(pred, dtype=F)
    B = pred.shape[0]
    delta, ratio = affine_params(kind, input_size, original_size)
    org = np.asarray(original_size, dtype=F).reshape(-1, 2)
    dvec = np.stack([delta[:, 1], delta[:, 0], delta[:, 1], delta[:, 0]], axis=-1)[:, None, :]
    with np.errstate(invalid="ignore", over="ignore"):
        coor = ((pred[..., 0:4] - dvec) / ratio[:, None, :]).astype(F)
        edge = (org - F(1.0))[:, ::-1][:, None, :]                      # (w-1, h-1)
        coor[..., 0:2] = np.maximum(coor[..., 0:2], F(0.0))
        coor[..., 2:4] = np.minimum(coor[..., 2:4], edge)
        score = (pred[..., 5:] * pred[..., 4:5]).astype(F)
    out = np.concatenate([np.broadcast_to(coor, (B,) + coor.shape[1:]), score], axis=-1)
    return np.ascontiguousarray(out, dtype=F)


# --------------------------------------------------------------------------------------------
# a6  threshold + class-aware NMS  (tools.py:540-566 + torchvision/ops/boxes.py:51-120)
# --------------------------------------------------------------------------------------------
def select_candidates(bboxes: np.ndarray, score_threshold: float):
    """tools.py:550-555: threshold the whole (N,C) score matrix in fp32; row-major nonzero."""
    bboxes = np.asarray(bboxes, dtype=F)
    scores = bboxes[:, 4:]
    rows, cls = np.nonzero(scores > F(score_threshold))
    return bboxes[rows, 0:4].copy(), scores[rows, cls].copy(), cls.astype(np.int64), rows.astype(np.int64)


def nms_plain(boxes, scores, iou_threshold: float, round_mode: int, use_mask: bool = False):
    """torchvision.ops.nms: stable descending order, greedy.  -> kept indices in visiting order."""
    boxes = np.ascontiguousarray(boxes, dtype=F)
    M = boxes.shape[0]
    if M == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-np.asarray(scores, dtype=F), kind="stable").astype(np.int64)
    keep_pos = np.empty((M,), np.int64)
    fn = _lib().pq_oracle_nms_mask if use_mask else _lib().pq_oracle_nms_ordered
    n = fn(_f32p(boxes), _i64p(order), M, float(iou_threshold), int(round_mode), _i64p(keep_pos))
    return order[keep_pos[:n]]


def batched_nms(boxes, scores, cls, iou_threshold: float, device: str = "cuda", mode: str = "auto"):
    """torchvision/ops/boxes.py:80-120.  device selects both the dispatch limit (cpu: numel > 4000
    -> vanilla, cuda: numel > 100000 -> vanilla) and the IoU rounding order (nms_oracle.c)."""
    boxes = np.ascontiguousarray(boxes, dtype=F)
    scores = np.ascontiguousarray(scores, dtype=F)
    cls = np.ascontiguousarray(cls, dtype=np.int64)
    M = boxes.shape[0]
    if M == 0:
        return np.zeros((0,), np.int64)
    round_mode = 1 if device == "cuda" else 0
    if mode == "auto":
        limit = 100_000 if device == "cuda" else 4000
        mode = "vanilla" if boxes.size > limit else "trick"
    if mode == "trick":
        shifted = np.empty_like(boxes)
        _lib().pq_oracle_trick_offsets(_f32p(boxes), _i64p(cls), M, _f32p(shifted))
        return nms_plain(shifted, scores, iou_threshold, round_mode)
    keep_mask = np.zeros((M,), bool)
    for c in np.unique(cls):
        idx = np.nonzero(cls == c)[0]
        k = nms_plain(boxes[idx], scores[idx], iou_threshold, round_mode)
        keep_mask[idx[k]] = True
    kept = np.nonzero(keep_mask)[0]
    return kept[np.argsort(-scores[kept], kind="stable")]


def torch_nms(bboxes: np.ndarray, score_threshold: float, iou_threshold: float,
              device: str = "cuda", mode: str = "auto", return_index: bool = False):
    """tools.py:540-566.  -> (K,6) [x1,y1,x2,y2,score,class]; shape (0,) when nothing is kept."""
    boxes, scores, cls, rows = select_candidates(bboxes, score_threshold)
    keep = batched_nms(boxes, scores, cls, iou_threshold, device=device, mode=mode)
    if keep.size == 0:
        out = np.zeros((0,), F)
    else:
        out = np.concatenate([boxes[keep], scores[keep, None], cls[keep, None].astype(F)], axis=1)
    if return_index:
        return out, rows[keep], cls[keep]
    return out


# --------------------------------------------------------------------------------------------
# a7/a8  IoU family  (tools.py:357-477)   broadcasting, last dim = (x1,y1,x2,y2)
# --------------------------------------------------------------------------------------------
def _parts(b1, b2):
    b1 = np.asarray(b1, dtype=F)
    b2 = np.asarray(b2, dtype=F)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    lu = np.maximum(b1[..., :2], b2[..., :2])
    rd = np.minimum(b1[..., 2:], b2[..., 2:])
    sec = np.maximum(rd - lu, F(0.0))
    inter = sec[..., 0] * sec[..., 1]
    union = (a1 + a2) - inter
    return b1, b2, inter, union


def iou_calc3(b1, b2):
    with np.errstate(invalid="ignore", divide="ignore"):
        _, _, inter, union = _parts(b1, b2)
        return (inter / union).astype(F)


def _giou_parts(b1, b2):
    b1, b2, inter, union = _parts(b1, b2)
    iou = inter / union
    elu = np.minimum(b1[..., :2], b2[..., :2])
    erd = np.maximum(b1[..., 2:], b2[..., 2:])
    enc = np.maximum(erd - elu, F(0.0))
    ea = enc[..., 0] * enc[..., 1]
    g = iou - (ea - union) / ea
    return b1, b2, iou, g, elu, erd


def giou(b1, b2):
    with np.errstate(invalid="ignore", divide="ignore"):
        return _giou_parts(b1, b2)[3].astype(F)


def diou(b1, b2):
    """tools.py:406-437 -- note the reference ADDS d_center/d_enclose."""
    with np.errstate(invalid="ignore", divide="ignore"):
        b1, b2, _, g, elu, erd = _giou_parts(b1, b2)
        c1 = (b1[..., :2] + b1[..., 2:]) / F(2.0)
        c2 = (b2[..., :2] + b2[..., 2:]) / F(2.0)
        dc = ((c1 - c2) ** 2).sum(axis=-1, dtype=F)
        de = ((elu - erd) ** 2).sum(axis=-1, dtype=F)
        return (g + dc / de).astype(F)


def ciou(b1, b2):
    """tools.py:439-477."""
    with np.errstate(invalid="ignore", divide="ignore"):
        b1, b2, iou, g, elu, erd = _giou_parts(b1, b2)
        w1, h1 = b1[..., 2] - b1[..., 0], b1[..., 3] - b1[..., 1]
        w2, h2 = b2[..., 2] - b2[..., 0], b2[..., 3] - b2[..., 1]
        c1 = (b1[..., :2] + b1[..., 2:]) / F(2.0)
        c2 = (b2[..., :2] + b2[..., 2:]) / F(2.0)
        dc = ((c1 - c2) ** 2).sum(axis=-1, dtype=F)
        de = ((elu - erd) ** 2).sum(axis=-1, dtype=F)
        v = F(4.0 / (np.pi ** 2)) * (np.arctan(w1 / h1) - np.arctan(w2 / h2)) ** 2
        alpha = v / ((F(1.0) - iou) + v)
        return (g + dc / de + alpha * v).astype(F)


def iou_calc1(b1, b2):
    """tools.py:335-355 (numpy, dtype-preserving, union clamped at 1e-14)."""
    b1 = np.asarray(b1)
    b2 = np.asarray(b2)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    lu = np.maximum(b1[..., :2], b2[..., :2])
    rd = np.minimum(b1[..., 2:], b2[..., 2:])
    sec = np.maximum(rd - lu, 0.0)
    inter = sec[..., 0] * sec[..., 1]
    return inter / np.maximum(a1 + a2 - inter, 1e-14)


def ignore_mask(pred_boxes: np.ndarray, gt: np.ndarray, thr: float) -> np.ndarray:
    """model/loss.py:85-90: (max_g iou_calc3(pred, gt_g)) < ignore_thresh, per prediction row.
    pred_boxes (R,4), gt (G,4) zero-padded.  C loop (nms_oracle.c) so that full-size inputs
    finish in seconds."""
    p = np.ascontiguousarray(pred_boxes, dtype=F).reshape(-1, 4)
    g = np.ascontiguousarray(gt, dtype=F).reshape(-1, 4)
    out = np.zeros((p.shape[0],), np.uint8)
    _lib().pq_oracle_ignore_mask(_f32p(p), p.shape[0], _f32p(g), g.shape[0], F(thr),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    return out.astype(bool)


# --------------------------------------------------------------------------------------------
# a11/a12  target assignment  (dataset/train_dataset.py:13-43, 109-150; tools.py:479-505)
# --------------------------------------------------------------------------------------------
def _iou_xywh_mixed(box_xywh_f32: np.ndarray, anchors_xywh_f64: np.ndarray) -> np.ndarray:
    """tools.py:479-505 with the dtypes create_label feeds it: box 1 is fp32 (area and corners
    computed in fp32), box 2 is fp64 (float32 // int64 promotes, train_dataset.py:132-134)."""
    b1 = np.asarray(box_xywh_f32, dtype=F)
    b2 = np.asarray(anchors_xywh_f64, dtype=np.float64)
    area1 = b1[2] * b1[3]                                   # fp32
    area2 = b2[:, 2] * b2[:, 3]                             # fp64
    c1 = np.concatenate([b1[:2] - b1[2:] * F(0.5), b1[:2] + b1[2:] * F(0.5)])            # fp32
    c2 = np.concatenate([b2[:, :2] - b2[:, 2:] * 0.5, b2[:, :2] + b2[:, 2:] * 0.5], axis=-1)
    lu = np.maximum(c1[:2].astype(np.float64), c2[:, :2])
    rd = np.minimum(c1[2:].astype(np.float64), c2[:, 2:])
    sec = np.maximum(rd - lu, 0.0)
    inter = sec[:, 0] * sec[:, 1]
    union = np.float64(area1) + area2 - inter
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / union


def create_label(bboxes: np.ndarray, output_sizes, num_classes: int, anchors,
                 strides=(8, 16, 32), iou_threshold: float = 0.3):
    """-> (label_s, label_m, label_l, sbboxes, mbboxes, lbboxes) for ONE image.

    bboxes (n,6) fp32 rows [x1,y1,x2,y2,class,mixw]; output_sizes (3,2) (h,w) per scale.
    Labels are zero with the last (mixw) channel 1.0; per GT: centre cell per scale, 9 anchor
    boxes centred on the cell centre, IoU > thr (else the argmax) selects (scale, ratio) slots;
    later GTs overwrite earlier ones; every hit appends the GT box to that scale's list.
    """
    strides_i = np.asarray(strides, dtype=np.int64)
    anchors = np.asarray(anchors, dtype=F)
    C = num_classes
    labels = []
    for i in range(3):
        lab = np.zeros((int(output_sizes[i][0]), int(output_sizes[i][1]), 3, 6 + C), dtype=F)
        lab[..., 6 + C - 1] = F(1.0)
        labels.append(lab)
    lists = [[], [], []]
    hot = 1.0 * (1 - 0.01) + 0.01 * (1.0 / C)              # fp64, stored as fp32 below
    cold = 0.0 * (1 - 0.01) + 0.01 * (1.0 / C)
    for box in np.asarray(bboxes, dtype=F).reshape(-1, 6):
        coor = box[:4]
        cls = int(box[4])
        xywh = np.concatenate([(coor[2:] + coor[:2]) * F(0.5), coor[2:] - coor[:2]]).astype(F)
        cell = np.floor_divide(xywh[:2].astype(np.float64)[:, None], strides_i.astype(np.float64))
        cell = cell.astype(np.int32).T                                       # (3 scales, 2) = (x, y)
        centre = (cell.astype(F) + F(0.5)).astype(np.float64) * strides_i[:, None]
        anc = np.concatenate([np.repeat(centre, 3, axis=0), anchors.astype(np.float64)], axis=-1)
        ious = _iou_xywh_mixed(xywh, anc)
        hit = ious > iou_threshold
        if not hit.any():
            hit[int(np.argmax(ious))] = True
        row = np.empty((6 + C,), dtype=np.float64)
        row[0:4] = coor
        row[4] = 1.0
        row[5:5 + C] = cold
        row[5 + cls] = hot
        row[5 + C] = box[5]
        for i in np.nonzero(hit)[0]:
            s, a = int(i) // 3, int(i) % 3
            x, y = int(cell[s][0]), int(cell[s][1])
            labels[s][y, x, a, :] = row.astype(F)
            lists[s].append(coor.copy())
    return labels[0], labels[1], labels[2], lists[0], lists[1], lists[2]


def collate_gt_lists(per_image_lists):
    """train_dataset.py:16-24: pad each image's (g,4) list with zero rows to the batch max (>= 1)."""
    n = max(max(len(l) for l in per_image_lists), 1)
    out = np.zeros((len(per_image_lists), n, 4), dtype=F)
    for b, l in enumerate(per_image_lists):
        if len(l):
            out[b, :len(l)] = np.asarray(l, dtype=F).reshape(-1, 4)
    return out


def create_label_batch(batch_bboxes, output_sizes, num_classes, anchors, strides=(8, 16, 32),
                       iou_threshold: float = 0.3):
    """create_label per image + collate_batch -> 3 label tensors (B,H,W,3,6+C), 3 GT tensors (B,G,4)."""
    per = [create_label(b, output_sizes, num_classes, anchors, strides, iou_threshold)
           for b in batch_bboxes]
    labels = [np.stack([p[i] for p in per], axis=0) for i in range(3)]
    gts = [collate_gt_lists([p[3 + i] for p in per]) for i in range(3)]
    return labels, gts


def iou_calc1(boxes1: np.ndarray, boxes2: np.ndarray) -> np.ndarray:
    """tools.py:335-355 (numpy, dtype of the inputs, union clamped at 1e-14)."""
    a1 = (boxes1[..., 2] - boxes1[..., 0]) * (boxes1[..., 3] - boxes1[..., 1])
    a2 = (boxes2[..., 2] - boxes2[..., 0]) * (boxes2[..., 3] - boxes2[..., 1])
    lu = np.maximum(boxes1[..., :2], boxes2[..., :2])
    rd = np.minimum(boxes1[..., 2:], boxes2[..., 2:])
    sec = np.maximum(rd - lu, 0.0)
    inter = sec[..., 0] * sec[..., 1]
    return inter / np.maximum(a1 + a2 - inter, 1e-14)


def numpy_nms(bboxes: np.ndarray, score_threshold: float, iou_threshold: float, sigma: float = 0.3,
              method: str = "nms") -> np.ndarray:
    """tools.py:507-538, class by class in ASCENDING class order (the reference iterates a Python set, whose order is
    unspecified): selection loop, hard NMS or soft-NMS decay, the first pick of a class not thresholded."""
    assert method in ("nms", "soft-nms")
    out = []
    for cls in sorted(set(bboxes[:, 5].tolist())):
        rows = bboxes[bboxes[:, 5] == cls].copy()
        while len(rows) > 0:
            m = int(np.argmax(rows[:, 4]))
            best = rows[m]
            out.append(best)
            rows = np.concatenate([rows[:m], rows[m + 1:]])
            iou = iou_calc1(best[np.newaxis, :4], rows[:, :4])
            weight = np.ones((len(iou),), dtype=np.float32)
            if method == "nms":
                weight[iou > iou_threshold] = 0.0
            else:
                weight = np.exp(-(1.0 * iou ** 2 / sigma))
            rows[:, 4] = rows[:, 4] * weight
            rows = rows[rows[:, 4] > score_threshold]
    return np.array(out)
