"""-m gpu: the evaluator statistics (pqdet_b200/evaluator.py -> pqdet_ap_match) against the reference's own
Evaluator.AP output (tests/golden/ap.npz) and against the oracle on larger / adversarial sets.  The (C,10) AP
table must be bit-identical: tp/fp flags are integers and the float arithmetic after them is the same numpy code."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ap_oracle
from pqdet_b200 import synth

pytestmark = pytest.mark.gpu


def _run(data, C):
    from pqdet_b200.evaluator import DetectionAccumulator
    acc = DetectionAccumulator(["c%d" % i for i in range(C)])
    orc = ap_oracle.ApOracle(C)
    for f, gt, diffs, dets in data:
        acc.add_detections(f, dets)
        acc.add_labels(f, gt, diffs)
        orc.add_detections(f, dets)
        orc.add_labels(f, gt, diffs)
    return acc, orc


def test_ap_golden_reference_evaluator():
    g = load_golden("ap")
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        n, C, size, seed = (int(v) for v in g["cfg_" + tag])
        acc, _ = _run(synth.make_eval_set(n, C, size, seed=seed, gt_dtype=dt), C)
        ap = acc.AP()
        assert np.array_equal(ap.raw, g["raw_" + tag]), tag
        assert np.array_equal(ap.mAPs, g["mAPs_" + tag]) and np.array_equal(ap.APs, g["APs_" + tag])
        assert float(ap.AP) == float(g["AP_" + tag])
        assert acc.detections_count == 0                       # AP() resets the statistics (eval/evaluator.py:138)


@pytest.mark.parametrize("dt,C,n,n_obj,seed", [(np.float32, 20, 300, (1, 12), 1), (np.float64, 10, 120, (20, 200), 2),
                                                (np.int64, 3, 60, (1, 30), 3), (np.float32, 80, 200, (0, 38), 4)])
def test_ap_vs_oracle(dt, C, n, n_obj, seed):
    lo, hi = n_obj
    data = synth.make_eval_set(n, C, 608, seed=seed, gt_dtype=dt, n_obj=(max(lo, 1), hi))
    if lo == 0:                                                # images whose detections have no label at all
        data = [(f, gt[:0], d[:0], dets) if i % 7 == 0 else (f, gt, d, dets) for i, (f, gt, d, dets) in enumerate(data)]
    acc, orc = _run(data, C)
    want = orc.AP()
    got = acc.AP()
    assert np.array_equal(got.raw, want)
    assert got.class_names == ["c%d" % i for i in range(C)] and len(got.iou_thresholds) == 10


def test_ap_edge_cases():
    from pqdet_b200.evaluator import DetectionAccumulator
    acc = DetectionAccumulator(["a", "b"])
    with pytest.raises(UnboundLocalError):                     # the reference fails the same way without detections
        acc.AP()
    # all ground truth difficult, identical boxes, NaN box, detection of a class without labels
    gt = np.array([[10, 10, 50, 50, 0], [10, 10, 50, 50, 0], [60, 60, 90, 90, 0]], np.float32)
    dets = np.array([[10, 10, 50, 50, 0.9, 0], [10, 10, 50, 50, 0.9, 0], [10, 10, 50, 50, 0.9, 0],
                     [np.nan, 10, 50, 50, 0.8, 0], [0, 0, 5, 5, 0.7, 1]], np.float32)
    for diffs in (np.array([1, 1, 1]), np.array([0, 1, 0]), np.array([0, 0, 0])):
        acc, orc = DetectionAccumulator(["a", "b"]), ap_oracle.ApOracle(2)
        for o in (acc, orc):
            o.add_detections("x", dets)
            o.add_labels("x", gt, diffs)
            o.add_detections("y", dets[:2])                    # image without any label
        with np.errstate(all="ignore"):
            want = orc.AP()
        got = acc.AP().raw
        assert np.array_equal(got, want, equal_nan=True), diffs


def test_ap_batch_ingest_from_fused_detections():
    """detections of a whole batch go from the fused kernel to the accumulator with one D2H copy."""
    from pqdet_b200 import fused
    from pqdet_b200.evaluator import DetectionAccumulator
    C, size, B = 20, 512, 8
    heads = [h.cuda() for h in synth.make_heads(B, C, size, "sparse", seed=9)]
    dets = fused.decode_nms(heads, synth.FPN_STRIDES, C, (size, size), torch.tensor([float(size)] * 2).cuda(), "voc")
    names = ["im%d" % i for i in range(B)]
    a, b = DetectionAccumulator(["c%d" % i for i in range(C)]), DetectionAccumulator(["c%d" % i for i in range(C)])
    a.add_detections_batch(names, dets)
    for i, n in enumerate(names):
        b.add_detections(n, dets[i].cpu().numpy())
    for i, n in enumerate(names):                               # labels = the top detections themselves
        top = dets[i][:5].cpu().numpy()
        for acc in (a, b):
            acc.add_labels(n, np.concatenate([top[:, :4], top[:, 5:6]], axis=1), np.zeros(len(top), np.int64))
    ra, rb = a.AP(), b.AP()
    # classes that were detected but never labelled have recall 0/0 = NaN, exactly like the reference
    assert np.array_equal(ra.raw, rb.raw, equal_nan=True) and float(np.nanmean(ra.raw)) > 0
