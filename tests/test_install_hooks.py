"""The monkey-patch hooks of pqdet_b200/install.py against the real reference tree (CPU only, in a
subprocess so the patched modules do not leak into the other tests).  Skipped where /root/reference is absent."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not present")


def test_install_patches_the_reference_hook_points():
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r)
        from oracle import ref_harness as rh
        ref = rh.load()
        import pqdet_b200.install as inst
        from pqdet_b200 import parser as pqp, tools as pqt, base_sample as pqb, loss as pql
        done = inst.install(strict=True, patch_augment=True)
        assert all(done.values()), done
        import tools, model.parser, model.loss, dataset, dataset.base_sample
        assert model.parser.YOLOLayer is pqp.YOLOLayer and model.parser.Decode is pqp.Decode
        assert model.loss.loss_per_scale is pql.loss_per_scale
        assert tools.torch_nms is pqt.torch_nms and tools.giou is pqt.giou
        assert dataset.RECOVER_BBOXES_REGISTER['voc'] is pqb.recover_bboxes_prediction_voc
        # the cfg parser now builds OUR layer for every [yolo] block, with the reference's opt dict
        from model.interpreter import DetectionModel
        m = DetectionModel(%r)
        yolo = [l for l in m.module_list if l._type == 'yolo']
        assert len(yolo) == 3 and all(isinstance(l, pqp.YOLOLayer) for l in yolo)
        assert [l.opt['stride'] for l in yolo] == [32, 16, 8]
        assert yolo[0].opt['classes'] == 20 and yolo[0].opt['bbox_loss'] == 'l1'
        # section 8f hooks: the evaluator's statistics and the eval letterbox
        import eval.evaluator, dataset.augment
        from pqdet_b200 import augment as pqa
        assert dataset.augment.Resize is pqa.Resize
        ev = eval.evaluator.Evaluator.__new__(eval.evaluator.Evaluator)
        ev._classes = ['a', 'b']
        ev.init_statics()
        assert type(ev._pq_acc).__name__ == 'DetectionAccumulator' and ev.detections_count == 0
        # section 8f-2 hook: the three head convolutions hand their input to the YOLOLayer in CUDA eval mode; nothing
        # about the module tree changes, and on CPU / in training mode the convolution still runs
        import torch
        keys = list(m.state_dict().keys())
        assert inst.fuse_head_convs(m) == 3 and list(m.state_dict().keys()) == keys
        head_conv = m.module_list[[i for i, l in enumerate(m.module_list) if l._type == 'yolo'][0] - 1]
        probe = torch.randn(1, head_conv.conv.in_channels, 2, 2)
        m.eval()
        assert tuple(head_conv(probe).shape) == (1, 75, 2, 2) and type(head_conv).__name__ == '_FusedHeadConv'
        import copy, pickle
        assert pickle.loads(pickle.dumps(head_conv)).conv.weight.shape == head_conv.conv.weight.shape
        # row a4 hook: the model's class gains a forward that combines the levels in one launch; module tree unchanged
        assert inst.fuse_eval_concat(m) and list(m.state_dict().keys()) == keys
        assert type(m).__mro__[1].__name__ == 'DetectionModel' and type(m).__mro__[2].__name__ == 'AnyModel'
        # training analogue: same subclass gains the one-launch loss route; on CPU the reference's own loop still runs
        assert inst.fuse_train_levels(m) and type(m)._pq_train_levels and type(m)._pq_eval_concat
        assert list(m.state_dict().keys()) == keys
        # the reference's own affine callables are recognised (dataset/*_sample.py), anything else is refused
        from dataset.voc_sample import _voc_affine_bboxes
        from dataset.coco_sample import _coco_affine_bboxes
        from dataset.visdrone_sample import _visdrone_affine_bboxes
        assert [pqb.affine_kind(f) for f in (_voc_affine_bboxes, _coco_affine_bboxes, _visdrone_affine_bboxes)] == \
            ['voc', 'coco', 'visdrone']
        try:
            pqb.affine_kind(lambda a, b: (a, b))
            raise SystemExit('custom affine accepted')
        except ValueError:
            pass
        # Evaluator.evaluate is replaced (one fused launch per batch); the loop keeps the reference's data protocol
        assert eval.evaluator.Evaluator.evaluate.__qualname__.startswith('_make_evaluate')
        assert done['eval.evaluator.Evaluator.evaluate']
        ev._score_threshold, ev._iou_threshold, ev._input_size = 0.1, 0.45, (64, 64)
        ev._recover_bboxes = dataset.RECOVER_BBOXES_REGISTER['voc']
        ev.model, ev.dataset = m, []
        assert inst._dataset_kind(ev._recover_bboxes) == 'voc'
        try:
            ev.evaluate()                                  # empty data set: the reference's AP() leaves `metrics` unbound
            raise SystemExit('empty evaluation returned')
        except UnboundLocalError:
            pass
        print('ok')
    """) % (ROOT, os.path.join(rh.REFERENCE_ROOT, "model", "cfg", "regnetx-600m-fpn.cfg"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=rh.REFERENCE_ROOT, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr
