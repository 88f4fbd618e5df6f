"""-m gpu: install.py's model-level hooks end to end on CUDA (tests/mini_darknet.py stands in for the cfg-built
reference model, which cannot travel to the GPU box): one launch per eval batch / per train step, same results as the
per-level route, autograd kept intact where someone differentiates through an eval prediction."""
import contextlib
import copy

import numpy as np
import pytest
import torch

from gpu_util import cuda
from mini_darknet import MiniDetectionModel
from oracle import pqdet_oracle as po

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def count_abi_calls():
    """Counts calls per C-ABI entry point while active (every entry point that launches is one or a few kernels)."""
    from pqdet_b200 import _lib
    lib = _lib.load()
    calls = {}
    originals = {}
    for name in _lib.SIGNATURES:
        fn = getattr(lib, name)
        originals[name] = fn

        def wrap(*a, _fn=fn, _name=name):
            calls[_name] = calls.get(_name, 0) + 1
            return _fn(*a)
        setattr(lib, name, wrap)
    try:
        yield calls
    finally:
        for name, fn in originals.items():
            setattr(lib, name, fn)


def _model(seed=0, **kw):
    torch.manual_seed(seed)
    m = MiniDetectionModel(**kw).cuda()
    for mod in m.modules():                       # spread the logits a little so that some boxes pass the threshold
        if isinstance(mod, torch.nn.Conv2d) and mod.bias is not None:
            torch.nn.init.normal_(mod.bias, 0.5, 1.5)
    return m


def test_eval_hooks_one_launch_and_same_rows():
    from pqdet_b200 import fused, install, tools, base_sample
    m = _model().eval()
    x = torch.randn(4, 3, 128, 160, device="cuda")
    with torch.no_grad():
        want = m(x)                                                    # per-level Decode + cat (3 launches + cat)
    assert install.fuse_eval_concat(m) and install.fuse_head_convs(m) == 3
    with torch.no_grad(), count_abi_calls() as calls:
        got = m(x)
    assert sum(calls.values()) <= 3 and set(calls) <= {"pqdet_head_conv_decode_levels", "pqdet_head_conv_decode"}
    # TF32 head conv vs cuDNN's (also TF32 by default, different accumulation order): close, not bit-identical
    assert got.shape == want.shape and torch.allclose(got, want, rtol=2e-2, atol=2e-2 * 160)
    # with TF32 off the hooks leave the convolution to PyTorch and the decode is then bit-identical
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad(), count_abi_calls() as calls:
            got32 = m(x)
            m2 = _model().eval()
            want32 = m2(x)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(got32, want32)
    assert calls.get("pqdet_decode_levels", 0) == 1 and "pqdet_head_conv_decode" not in calls
    # raw heads -> fused decode+NMS == prediction -> recover -> per-image torch_nms
    raws, strides, C = install.raw_heads(m, x)
    assert strides == [32, 16, 8] and C == 4 and [tuple(r.shape[2:]) for r in raws] == [(4, 5), (8, 10), (16, 20)]
    orig = torch.tensor([[100., 150.], [128., 160.], [90., 160.], [128., 100.]], device="cuda")
    with count_abi_calls() as calls:
        dets = fused.decode_nms(raws, strides, C, (128, 160), orig, "voc", 0.7, 0.45)
        rows = dets.to_numpy_list()
    assert calls == {"pqdet_decode_nms": 1}
    dec = torch.cat([l.decode(r).view(4, -1, 5 + C) for l, r in
                     zip([l for l in m.module_list if l._type == 'yolo'], raws)], dim=1)
    rec = base_sample.recover_bboxes_prediction_voc(dec, (128, 160), orig)
    assert sum(len(r) for r in rows) > 0
    for b in range(4):
        w = tools.torch_nms(rec[b], 0.7, 0.45).cpu().numpy().reshape(-1, 6)
        assert np.array_equal(rows[b].reshape(-1, 6), w)


def test_evaluate_hook_runs_one_fused_launch_per_batch():
    """Evaluator.evaluate's body (eval/evaluator.py:47-61) through install.detect_batch / _make_evaluate on an
    object with the fields the reference Evaluator has."""
    from pqdet_b200 import base_sample, install, tools
    from pqdet_b200.evaluator import DetectionAccumulator
    m = _model(1).eval()
    assert install.fuse_eval_concat(m)

    class Ev:
        evaluate = install._make_evaluate(None)
        predict = lambda self, imgs: self.model(imgs)

        def __init__(self, model, dataset):
            self.model, self.dataset = model, dataset
            self._score_threshold, self._iou_threshold, self._input_size = 0.7, 0.45, (128, 128)
            self._recover_bboxes = base_sample.RECOVER_BBOXES_REGISTER['coco']
            self.acc = DetectionAccumulator(['a', 'b', 'c', 'd'])
            self.seen = []

        def add_detections(self, f, bboxes):
            self.seen.append((f, np.array(bboxes)))
            self.acc.add_detections(f, bboxes)

        def add_labels(self, f, labels, diffs):
            self.acc.add_labels(f, labels, diffs)

        def AP(self):
            return self.acc.AP()

    g = torch.Generator().manual_seed(0)
    batches = []
    for k in range(2):
        imgs = torch.randn(3, 3, 128, 128, generator=g).cuda()
        names = ["img%d_%d" % (k, i) for i in range(3)]
        shapes = torch.tensor([[100., 128.], [128., 128.], [128., 90.]])
        labels = [np.array([[10., 10., 60., 60., i % 4]], np.float32) for i in range(3)]
        diffs = [np.zeros((1,), bool) for _ in range(3)]
        batches.append((imgs, names, shapes, labels, diffs))
    ev = Ev(m, batches)
    with count_abi_calls() as calls:
        ap = ev.evaluate()
    assert calls.get("pqdet_decode_nms", 0) == 2 and "pqdet_decode_levels" not in calls and "pqdet_recover" not in calls
    assert len(ev.seen) == 6 and ap.raw.shape == (4, 10)
    # same rows as the reference's loop body on the unfused route
    with torch.no_grad():
        for k, (imgs, names, shapes, _, _) in enumerate(batches):
            rec = ev._recover_bboxes(m(imgs), torch.tensor([128., 128.]).cuda(), shapes.cuda())
            for i in range(3):
                w = tools.torch_nms(rec[i], 0.7, 0.45).cpu().numpy()
                got = ev.seen[3 * k + i][1]
                assert got.shape == w.shape and np.array_equal(got, w), (k, i)
    # with the head convolutions deferred too (fuse_head_convs) the route starts at the conv inputs: one conv launch
    # per level whose epilogue thresholds + one NMS launch on the hit records; same rows up to TF32 vs cuDNN rounding,
    # so compare against the same kernels' raw-head route
    m3 = _model(1).eval()
    assert install.fuse_eval_concat(m3) and install.fuse_head_convs(m3) == 3
    ev3 = Ev(m3, batches)
    with count_abi_calls() as calls:
        ev3.evaluate()
    if "pqdet_records_nms" in calls:          # 128 x 128 inputs: 16 / 64 / 256 cells per level - only levels of >= 128 cells qualify
        assert calls["pqdet_records_nms"] == 2 and calls["pqdet_head_conv_hits"] == 6
    else:
        assert calls.get("pqdet_decode_nms", 0) == 2 and calls.get("pqdet_head_conv_decode", 0) == 6
    assert len(ev3.seen) == 6
    # a model without the hook takes predict -> recover -> ONE batched NMS launch
    ev2 = Ev(_model(1).eval(), batches)
    with count_abi_calls() as calls:
        ev2.evaluate()
    assert calls.get("pqdet_nms_fused", 0) == 2 and calls.get("pqdet_recover", 0) == 2
    for a, b in zip(ev.seen, ev2.seen):
        assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("kind", ["giou", "l1"])
def test_train_hook_one_launch_same_dict_and_gradients(kind):
    from pqdet_b200 import install, synth
    from pqdet_b200.train_dataset import LabelAssigner
    B, C, size = 4, 4, 128
    m1 = _model(2, bbox_loss=kind).train()
    m2 = copy.deepcopy(m1)
    assert install.fuse_train_levels(m2)
    gts = synth.make_gt(B, C, size, 1, 6, seed=3)
    out_sizes = np.array([[size // 8] * 2, [size // 16] * 2, [size // 32] * 2])
    target = LabelAssigner(C).create_label_batch(gts, out_sizes)
    x = torch.randn(B, 3, size, size, device="cuda")
    with count_abi_calls() as c1:
        o1 = m1(x, target)
        o1['loss'].mean().backward()
    with count_abi_calls() as c2:
        o2 = m2(x, target)
        o2['loss'].mean().backward()
    assert c1.get("pqdet_loss_fwd_bwd", 0) == 3                       # the reference's route: one call per level
    assert c2.get("pqdet_loss_levels", 0) == 1 and "pqdet_loss_fwd_bwd" not in c2
    assert set(o2) == set(o1) and len(o2['loss_per_branch']) == 3
    for k in ("loss", "giou_loss", "conf_loss", "class_loss"):
        assert tuple(o2[k].shape) == (1,)
        assert abs(float(o2[k]) - float(o1[k])) <= 1e-5 * abs(float(o1[k])), k
    for a, b in zip(o1['loss_per_branch'], o2['loss_per_branch']):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(a))
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2 and p1.grad is not None and p2.grad is not None
        scale = float(p1.grad.abs().max()) + 1e-12
        assert float((p1.grad - p2.grad).abs().max()) <= 1e-4 * scale, n1
    # eval-mode validation loss (target given, model.eval(), head convs deferred by fuse_head_convs)
    assert install.fuse_head_convs(m2) == 3 and install.fuse_eval_concat(m2)
    m1.eval(); m2.eval()
    with torch.no_grad():
        v1, v2 = m1(x, target), m2(x, target)
    assert abs(float(v1['loss']) - float(v2['loss'])) <= 1e-5 * abs(float(v1['loss']))


def test_eval_mode_backward_through_fused_model_matches_unfused():
    """saliency / adversarial / distillation code differentiates through an eval-mode prediction; the
    hooks must not hand back a detached tensor."""
    from pqdet_b200 import install
    m1 = _model(5).eval()
    m2 = copy.deepcopy(m1)
    assert install.fuse_eval_concat(m2) and install.fuse_head_convs(m2) == 3
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # both sides full-precision convs: results comparable bit for bit
    try:
        x1 = torch.randn(2, 3, 128, 128, device="cuda").requires_grad_(True)
        x2 = x1.detach().clone().requires_grad_(True)
        p1, p2 = m1(x1), m2(x2)
        assert p2.grad_fn is not None
        assert torch.equal(p1, p2)
        p1[..., 4].sum().backward()
        p2[..., 4].sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert x2.grad is not None and torch.allclose(x1.grad, x2.grad, rtol=1e-5, atol=1e-7)
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert (a.grad is None) == (b.grad is None)
        if a.grad is not None:
            assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7)
    # frozen parameters + input without grad + grad mode on: nothing to differentiate -> the fast route
    for p in m2.parameters():
        p.requires_grad_(False)
    assert m2(x2.detach()).grad_fn is None


def test_recover_accepts_affines_by_name_and_rejects_others():
    from pqdet_b200 import base_sample

    def _coco_affine_bboxes(input_size, batch_original_size):          # what the reference's module defines
        raise AssertionError("never called: the kernel computes the affine")

    pred = torch.rand(2, 50, 9, device="cuda") * 100
    orig = torch.tensor([[80., 100.], [100., 60.]], device="cuda")
    a = base_sample.recover_bboxes_prediction(pred, (128, 128), orig, _coco_affine_bboxes)
    b = base_sample.recover_bboxes_prediction_coco(pred, (128, 128), orig)
    c = base_sample.recover_bboxes_prediction(pred, (128, 128), orig, "coco")
    assert torch.equal(a, b) and torch.equal(a, c)
    with pytest.raises(ValueError):
        base_sample.recover_bboxes_prediction(pred, (128, 128), orig, lambda i, o: (0, 1))
