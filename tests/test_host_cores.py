"""CPU check of the arithmetic cores shared by every CUDA kernel (pqdet_b200/csrc/pq_math.cuh),
compiled as host C++ by tests/host_harness/harness.cpp.  Catches formula errors without a GPU;
the -m gpu tests remain the parity tests proper.  Not a product path: nothing in pqdet_b200 loads it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, rel_close
from oracle import loss_ref
from oracle import pqdet_oracle as po

F = ctypes.POINTER(ctypes.c_float)
I = ctypes.POINTER(ctypes.c_int)


def fp(a):
    return a.ctypes.data_as(F)


@pytest.fixture(scope="module")
def hh(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hh") / "libhh.so")
    src = os.path.join(ROOT, "tests", "host_harness", "harness.cpp")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c++17", "-o", out, src, "-lm"])
    return ctypes.CDLL(out)


def test_decode_and_recover_cores(hh):
    g = load_golden("decode")
    C = int(g["num_classes"])
    for s in (32, 16, 8):
        raw = np.ascontiguousarray(g["raw_s%d" % s])
        B, CH, H, W = raw.shape
        out = np.empty((B, H, W, 3, 5 + C), np.float32)
        hh.hh_decode(fp(raw), fp(out), B, 3, C, H, W, ctypes.c_float(s))
        want = g["out_s%d" % s]
        assert rel_close(out[..., :4], want[..., :4], 1e-5, scale=float(s * max(H, W)))
        assert rel_close(out[..., 4:], want[..., 4:], 1e-5, scale=1e-30)
    r = load_golden("recover")
    pred = np.ascontiguousarray(r["pred"])
    B, N, ch = pred.shape
    for kind, k in (("voc", 0), ("coco", 1), ("visdrone", 2)):
        out = np.empty((B, N, ch - 1), np.float32)
        orig = np.ascontiguousarray(r["orig"])
        hh.hh_recover(fp(pred), fp(out), B, ctypes.c_int64(N), ch - 5, k, ctypes.c_float(r["input_size"][0]),
                      ctypes.c_float(r["input_size"][1]), fp(orig), 1)
        assert np.array_equal(out, r["out_" + kind]), kind


def test_nms_pair_core_matches_c_oracle(hh):
    rng = np.random.default_rng(0)
    c = rng.random((4000, 2), np.float32) * 50
    wh = rng.random((4000, 2), np.float32) * 30 + 1
    boxes = np.concatenate([c - wh / 2, c + wh / 2], axis=1).astype(np.float32)
    hh.hh_nms_suppresses.argtypes = [F, F, ctypes.c_double, ctypes.c_int]
    n_dis = 0
    for i in range(0, 3998, 2):
        a, b = np.ascontiguousarray(boxes[i]), np.ascontiguousarray(boxes[i + 1])
        for thr in (0.45, 0.65):
            res = [hh.hh_nms_suppresses(fp(a), fp(b), thr, rnd) for rnd in (0, 1)]
            # reference formulas spelled out in numpy
            l, t = max(a[0], b[0]), max(a[1], b[1])
            r, d = min(a[2], b[2]), min(a[3], b[3])
            w, h = max(np.float32(r - l), np.float32(0)), max(np.float32(d - t), np.float32(0))
            inter = np.float32(w * h)
            Sa = np.float32(np.float32(a[2] - a[0]) * np.float32(a[3] - a[1]))
            bw, bh = np.float32(b[2] - b[0]), np.float32(b[3] - b[1])
            fma = np.float32(np.float64(bw) * np.float64(bh) + np.float64(Sa))       # exact product, one rounding
            cuda_q = np.float32(inter / np.float32(fma - inter)) > np.float32(thr)
            cpu_q = float(np.float32(inter / np.float32(np.float32(Sa + np.float32(bw * bh)) - inter))) > thr
            assert res[0] == int(cuda_q) and res[1] == int(cpu_q)
            n_dis += int(cuda_q != cpu_q)
    assert n_dis >= 0


def test_nms_pair_core_borderline_quotients(hh):
    """Walk the shift of box b float by float across the point where I/D crosses thr and compare every step with
    the reference formulas (both rounding orders): the pair test must agree on every borderline quotient."""
    hh.hh_nms_suppresses.argtypes = [F, F, ctypes.c_double, ctypes.c_int]
    n_flip = 0
    for (W, H, y) in ((37.25, 21.5, 0.0), (103.0, 64.75, 3.5), (9.125, 300.0, -2.25)):
        for thr in (0.45, 0.65, 0.5, 0.1, 1.0 / 3.0):
            a = np.array([10.0, 5.0, 10.0 + W, 5.0 + H], np.float32)
            hb = H - abs(y)
            # IoU(x) = (W-x)*hb / (2*W*H - (W-x)*hb) = thr  ->  x analytically, then +-150 floats around it
            inter = thr * 2 * W * H / (1 + thr)
            x0 = np.float32(W - inter / hb)
            x = x0
            for _ in range(150):
                x = np.nextafter(x, np.float32(-1e9))
            seen = set()
            for _ in range(300):
                b = np.array([10.0 + x, 5.0 + y, 10.0 + x + W, 5.0 + y + H], np.float32)
                l, t = max(a[0], b[0]), max(a[1], b[1])
                r, d = min(a[2], b[2]), min(a[3], b[3])
                w, h = max(np.float32(r - l), np.float32(0)), max(np.float32(d - t), np.float32(0))
                I = np.float32(w * h)
                Sa = np.float32(np.float32(a[2] - a[0]) * np.float32(a[3] - a[1]))
                bw, bh = np.float32(b[2] - b[0]), np.float32(b[3] - b[1])
                fma = np.float32(np.float64(bw) * np.float64(bh) + np.float64(Sa))
                cuda_q = bool(np.float32(I / np.float32(fma - I)) > np.float32(thr))
                cpu_q = bool(float(np.float32(I / np.float32(np.float32(Sa + np.float32(bw * bh)) - I))) > thr)
                got = [hh.hh_nms_suppresses(fp(a), fp(np.ascontiguousarray(b)), thr, rnd) for rnd in (0, 1)]
                assert got == [int(cuda_q), int(cpu_q)], (W, H, y, thr, float(x))
                seen.add(cuda_q)
                x = np.nextafter(x, np.float32(1e9))
            n_flip += int(len(seen) == 2)
    assert n_flip >= 10          # the walks really straddle the threshold


def test_iou_family_core(hh):
    g = load_golden("iou")
    b1, b2 = np.ascontiguousarray(g["b1"]), np.ascontiguousarray(g["b2"])
    n = b1.shape[0]
    for kind, name in enumerate(("iou_calc3", "giou", "diou", "ciou")):
        out = np.empty((n,), np.float32)
        hh.hh_iou(fp(b1), fp(b2), fp(out), ctypes.c_int64(n), kind)
        assert rel_close(out, g[name], 1e-5, scale=1.0), name
    for kind, fn in enumerate((loss_ref.iou_t, loss_ref.giou_t, lambda p, q: loss_ref.giou_t(p, q, True))):
        p = torch.from_numpy(b1).requires_grad_(True)
        q = torch.from_numpy(b2).requires_grad_(True)
        fn(p, q).sum().backward()
        g1, g2 = np.empty_like(b1), np.empty_like(b2)
        hh.hh_iou_grad(fp(b1), fp(b2), fp(g1), fp(g2), ctypes.c_int64(n), kind)
        assert rel_close(g1, p.grad.numpy(), 1e-4, scale=float(p.grad.abs().max())), kind
        assert rel_close(g2, q.grad.numpy(), 1e-4, scale=float(q.grad.abs().max())), kind


@pytest.mark.parametrize("kind", ["l1", "iou", "giou", "diou"])
def test_loss_core_golden(hh, kind):
    g = load_golden("train")
    C = int(g["num_classes"])
    for s in (8, 16, 32):
        raw = np.ascontiguousarray(g["raw_s%d" % s])
        lab = np.ascontiguousarray(g["label_s%d" % s])
        gt = np.ascontiguousarray(g["gtlist_s%d" % s])
        B, _, H, W = raw.shape
        grad = np.empty_like(raw)
        out4 = np.zeros((4,), np.float64)
        hh.hh_loss(fp(raw), fp(lab), fp(gt), fp(grad), out4.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                   B, 3, C, H, W, gt.shape[1], ctypes.c_float(s), ["l1", "iou", "giou", "diou"].index(kind),
                   ctypes.c_float(0.5), ctypes.c_float(0.05))
        assert rel_close(out4, g["loss_%s_s%d" % (kind, s)], 1e-5, scale=1e-30), (kind, s, out4)
        gg = g["grad_%s_s%d" % (kind, s)]
        assert rel_close(grad, gg, 1e-5, scale=float(np.abs(gg).max())), (kind, s)


def test_assign_core_golden_and_random(hh):
    from pqdet_b200 import synth
    g = load_golden("train")
    C, size = int(g["num_classes"]), int(g["size"])
    cases = [(C, size, [g["gt"][b, :int(n)] for b, n in enumerate(g["gt_counts"])], g["anchors"])]
    vis = np.array([(9, 13), (25, 17), (16, 31), (47, 29), (32, 51), (83, 48), (61, 91), (131, 99), (210, 189)], np.float32)
    cases.append((10, 608, synth.make_gt(3, 10, 608, 20, 200, seed=1), vis))
    for C, size, gts, anchors in cases:
        Hs = (ctypes.c_int * 3)(size // 8, size // 16, size // 32)
        st = (ctypes.c_int * 3)(8, 16, 32)
        out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
        anc = np.ascontiguousarray(anchors, np.float32)
        for gb in gts:
            gb = np.ascontiguousarray(gb, np.float32)
            labels = [np.empty((size // s, size // s, 3, 6 + C), np.float32) for s in (8, 16, 32)]
            cap = max(3 * len(gb), 1)
            lists = np.zeros((3, cap, 4), np.float32)
            ll = (ctypes.c_int * 3)()
            hh.hh_assign(fp(gb), len(gb), C, fp(anc), st, Hs, Hs, ctypes.c_double(np.float32(0.3)), fp(labels[0]),
                         fp(labels[1]), fp(labels[2]), fp(lists), cap, ll)
            want = po.create_label(gb, out_sizes, C, anc)
            for i in range(3):
                assert np.array_equal(labels[i], want[i])
                wl = np.asarray(want[3 + i], np.float32).reshape(-1, 4)
                assert ll[i] == len(wl) and np.array_equal(lists[i, :ll[i]], wl)
