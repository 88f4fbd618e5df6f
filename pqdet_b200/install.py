"""Monkey-patch the hot path into an importable copy of the reference (never edits its files).

    import tools                      # the reference's modules, imported by the user as usual
    import pqdet_b200.install
    pqdet_b200.install.install()

After this, DetectionModel built from any darknet cfg decodes / computes its loss with the sm_100a
kernels, Evaluator / predict.py recover boxes and run NMS with them, and TrainDataset.create_label
is available on the GPU.  The hook points are the ones SURVEY.md section 8b lists.
"""
from __future__ import annotations

import importlib
import sys

import torch
from torch import nn

from . import base_sample, loss, parser, tools as pq_tools


def install(strict: bool = False, patch_augment: bool = False, patch_evaluate: bool = True) -> dict:
    """Patch every reference module that is importable; returns {attribute path: True/False}.
    patch_augment: also replace dataset.augment.Resize by the GPU letterbox.  Off by default: DataLoader worker
    processes (where the training pipeline calls Resize) must not touch CUDA; turn it on for single-process eval /
    predict scripts, or call pqdet_b200.augment.letterbox_normalize on whole batches instead.
    patch_evaluate: replace Evaluator.evaluate's per-image loop (eval/evaluator.py:52-61) by one fused launch per batch
    (see detect_batch); the model must have been prepared with fuse_eval_concat to take the raw-head route."""
    done = {}

    def patch(modname, attr, value):
        key = "%s.%s" % (modname, attr)
        try:
            mod = sys.modules.get(modname) or importlib.import_module(modname)
            setattr(mod, attr, value)
            done[key] = True
        except Exception:
            if strict:
                raise
            done[key] = False

    for name in ("torch_nms", "iou_calc3", "giou", "diou", "ciou"):
        patch("tools", name, getattr(pq_tools, name))
    patch("model.loss", "loss_per_scale", loss.loss_per_scale)
    patch("model.parser", "Decode", parser.Decode)
    patch("model.parser", "YOLOLayer", parser.YOLOLayer)
    patch("model.parser", "loss_per_scale", loss.loss_per_scale)
    patch("dataset.base_sample", "recover_bboxes_prediction", base_sample.recover_bboxes_prediction)
    for ds in ("voc", "coco", "visdrone"):
        fn = base_sample.RECOVER_BBOXES_REGISTER[ds]
        patch("dataset.%s_sample" % ds, "recover_bboxes_prediction_%s" % ds, fn)
        try:
            importlib.import_module("dataset").RECOVER_BBOXES_REGISTER[ds] = fn
            done["dataset.RECOVER_BBOXES_REGISTER[%s]" % ds] = True
        except Exception:
            if strict:
                raise
            done["dataset.RECOVER_BBOXES_REGISTER[%s]" % ds] = False
    # the steps either side of the path: evaluator statistics and the eval letterbox
    try:
        _patch_evaluator(patch_evaluate)
        if patch_evaluate:
            done["eval.evaluator.Evaluator.evaluate"] = True
        done["eval.evaluator.Evaluator.{init_statics,add_detections,add_labels,AP}"] = True
    except Exception:
        if strict:
            raise
        done["eval.evaluator.Evaluator.{init_statics,add_detections,add_labels,AP}"] = False
    if patch_augment:
        from . import augment as pq_augment
        patch("dataset.augment", "Resize", pq_augment.Resize)
    return done


class _FusedHeadConv(nn.Sequential):
    """Class a head-convolution block is switched to by fuse_head_convs (same attributes, same state_dict keys).  In
    CUDA eval mode it does not convolve: it passes its input on, tagged with its own nn.Conv2d, and the YOLOLayer that
    receives the tagged tensor runs convolution + Decode as one kernel.  No reference to another module is kept, so
    nn.DataParallel replicas (tools.py:215-216) work on their own parameters."""

    def forward(self, x):
        if self.training or not x.is_cuda or torch.onnx.is_in_onnx_export():
            return super().forward(x)
        if torch.is_grad_enabled() and (x.requires_grad or self.conv.weight.requires_grad
                                        or (self.conv.bias is not None and self.conv.bias.requires_grad)):
            return super().forward(x)         # someone differentiates through the prediction: keep autograd intact
        if not (torch.backends.cudnn.allow_tf32 and torch.backends.cudnn.enabled):
            return super().forward(x)         # the tensor-core kernel is TF32; the user asked for full fp32 convs
        y = x.view_as(x)                      # a new tensor object on the same storage: the tag stays private
        y._pq_pending_conv = self.conv
        return y


def fuse_head_convs(model) -> int:
    """SURVEY 8f-2 as a hook: let every `[yolo]` level of a reference DetectionModel (built after install()) run its
    1x1 head convolution + Decode as ONE tensor-core kernel in eval mode.

    For each pqdet_b200.parser.YOLOLayer in `model.module_list` whose predecessor is the plain head convolution of the
    cfgs (`filters=A*(5+C), size=1, stride=1, activation=linear`, no batch norm: an nn.Sequential holding only `conv`,
    model/parser.py:385-410) and whose output no route / shortcut reads, the conv block's class is switched to
    _FusedHeadConv: in eval mode on CUDA it hands its INPUT to the YOLOLayer together with its own parameters
    (pqdet_head_conv_decode); in training mode, on CPU, for ONNX export or when a target is given the convolution
    still runs - and also whenever autograd is recording through the level (grad mode on and the input or the conv's
    parameters require grad: the fused kernel has no backward) or the user turned TF32 convolutions off
    (torch.backends.cudnn.allow_tf32 = False; the tensor-core kernel multiplies in TF32 like cuDNN's default).
    No module is added or renamed, so state_dict keys, pruning and checkpoint loading are unchanged.
    Returns the number of levels fused."""
    layers = list(model.module_list)
    used = set()
    for j, layer in enumerate(layers):              # who reads which cached output (model/interpreter.py:46-50)
        t = getattr(layer, '_type', None)
        refs = []
        if t in ('shortcut', 'scale_channels'):
            refs = [layer._from]
        elif t == 'route':
            refs = list(layer._layers)
        for r in refs:
            used.add(r if r >= 0 else j + r)
    fused = 0
    for i, layer in enumerate(layers):
        if i == 0 or not isinstance(layer, parser.YOLOLayer) or getattr(model, 'quant', False):
            continue
        prev = layers[i - 1]
        if isinstance(prev, _FusedHeadConv):
            fused += 1
            continue
        conv = getattr(prev, 'conv', None)
        if (getattr(prev, '_type', None) != 'convolutional' or type(prev) is not nn.Sequential
                or not isinstance(conv, nn.Conv2d) or len(prev) != 1):
            continue
        if (conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.padding != (0, 0) or conv.dilation != (1, 1)
                or conv.groups != 1 or conv.out_channels % (5 + layer.opt['classes']) or (i - 1) in used):
            continue
        prev.__class__ = _FusedHeadConv
        fused += 1
    return fused


def _fuse_model_forward(model, eval_concat: bool, train_levels: bool) -> bool:
    """Switch the model's class to a subclass whose forward runs the parent's layer loop (AnyModel.forward) with the
    YOLOLayers passing their input through, and then combines the levels itself (interpreter.combine_eval /
    combine_train).  Flags accumulate over calls.  Returns False (model untouched) if a [yolo] layer is not ours or
    something reads a [yolo] output."""
    from . import interpreter as pq_interpreter
    cls = type(model)
    if getattr(cls, '_pq_fused_forward', False):
        cls._pq_eval_concat = cls._pq_eval_concat or eval_concat
        cls._pq_train_levels = cls._pq_train_levels or train_levels
        return True
    layers = list(model.module_list)
    yolo_idx = [i for i, l in enumerate(layers) if getattr(l, '_type', None) == 'yolo']
    if not yolo_idx or not all(isinstance(layers[i], parser.YOLOLayer) for i in yolo_idx) or getattr(model, 'quant', False):
        return False
    for j, layer in enumerate(layers):
        t = getattr(layer, '_type', None)
        refs = [layer._from] if t in ('shortcut', 'scale_channels') else list(layer._layers) if t == 'route' else []
        if any((r if r >= 0 else j + r) in yolo_idx for r in refs):
            return False
    loop = None
    for k in cls.__mro__[1:]:
        if 'forward' in k.__dict__:
            loop = k.__dict__['forward']          # AnyModel.forward: the layer loop, returns the [yolo] outputs
            break
    if loop is None:
        return False

    def collect(self, x, target=None):
        """-> (YOLOLayers in cfg order, what each of them received)."""
        yolos = [l for l in self.module_list if isinstance(l, parser.YOLOLayer)]
        for l in yolos:
            object.__setattr__(l, '_pq_passthrough', True)
        try:
            outs = loop(self, x, target)
        finally:
            for l in yolos:
                l.__dict__.pop('_pq_passthrough', None)
        if not isinstance(outs, (list, tuple)):
            outs = [outs]
        return yolos, list(outs)

    def forward(self, x, target=None):
        me = type(self)
        if not x.is_cuda or torch.onnx.is_in_onnx_export():
            return cls.forward(self, x, target)
        if target is None:
            if not me._pq_eval_concat or self.training:
                return cls.forward(self, x, None)
            return pq_interpreter.combine_eval(*collect(self, x))
        if not me._pq_train_levels:
            return cls.forward(self, x, target)
        yolos, outs = collect(self, x, target)
        return pq_interpreter.combine_train(yolos, outs, target)

    model.__class__ = type(cls.__name__, (cls,), {
        'forward': forward, '_pq_collect': collect, '_pq_fused_forward': True,
        '_pq_eval_concat': eval_concat, '_pq_train_levels': train_levels})
    return True


def fuse_eval_concat(model) -> bool:
    """SURVEY section 8 row a4 as a hook: the eval branch of DetectionModel.forward (model/interpreter.py:72-76: every
    [yolo] level decoded, viewed as (B, -1, 5+C) and concatenated) becomes one launch that decodes all levels
    straight into the (B, N, 5+C) prediction - no per-level tensor, no torch.cat (which alone moves as many bytes as
    the decode).  Combined with fuse_head_convs the head convolutions run inside that launch too.

    Training mode, CPU tensors and ONNX export take the reference's own forward; when autograd is recording through
    the prediction (eval mode, grad enabled, inputs requiring grad) the levels go through the autograd-aware Decode
    and torch.cat, so the result keeps its grad_fn."""
    return _fuse_model_forward(model, True, False)


def fuse_train_levels(model) -> bool:
    """The training analogue: DetectionModel.forward(x, target) (model/interpreter.py:77-85: one YOLOLayer call per
    level, then Python sums over the per-level tuples) becomes ONE launch of the multi-level decode + loss + gradient
    kernel (pqdet_loss_levels) behind one autograd node; the dict it returns has the reference's keys and shapes.
    Levels with different loss options fall back to one launch per level."""
    return _fuse_model_forward(model, False, True)


def raw_heads(model, x):
    """Run a model prepared by fuse_eval_concat / fuse_train_levels up to its [yolo] layers.
    -> (raw head tensors in cfg order, strides, num_classes).  Head convolutions deferred by fuse_head_convs are
    applied here.  This is what the fused decode + NMS kernel consumes (fused.decode_nms)."""
    cls = type(model)
    if not getattr(cls, '_pq_fused_forward', False):
        raise ValueError("call install.fuse_eval_concat(model) first")
    yolos, outs = cls._pq_collect(model, x)
    raws = []
    for o in outs:
        conv = getattr(o, '_pq_pending_conv', None)
        raws.append(o if conv is None else torch.nn.functional.conv2d(o, conv.weight, conv.bias))
    return raws, [l.opt['stride'] for l in yolos], yolos[0].opt['classes']


def _patch_evaluator(patch_evaluate: bool = True):
    """Evaluator's statistics go through DetectionAccumulator, so AP() runs the matching on the GPU and returns the
    same tools.AP tuple; with patch_evaluate its per-batch body (eval/evaluator.py:47-61: predict, recover, then per
    image torch_nms + .cpu() + add_detections) becomes one fused launch and one device->host copy per batch."""
    ev_mod = importlib.import_module("eval.evaluator")
    ref_tools = sys.modules.get("tools") or importlib.import_module("tools")
    from .evaluator import DetectionAccumulator

    def init_statics(self):
        self._pq_acc = DetectionAccumulator(self._classes)
        self.detections_count = 0

    def add_detections(self, file_name, bboxes):
        self._pq_acc.add_detections(file_name, bboxes)
        self.detections_count = self._pq_acc.detections_count

    def add_labels(self, file_name, bboxes, diffs):
        self._pq_acc.add_labels(file_name, bboxes, diffs)

    def AP(self):
        m = self._pq_acc.AP()
        self.init_statics()
        return ref_tools.AP(m.mAPs, m.APs, m.AP, m.raw, m.class_names, m.iou_thresholds)

    for name, fn in (("init_statics", init_statics), ("add_detections", add_detections),
                     ("add_labels", add_labels), ("AP", AP)):
        setattr(ev_mod.Evaluator, name, fn)
    if patch_evaluate:
        ev_mod.Evaluator.evaluate = _make_evaluate(getattr(ev_mod, "tqdm", None))
        ev_mod.Evaluator.pq_detect_batch = detect_batch


def _dataset_kind(recover_fn) -> str:
    name = getattr(recover_fn, "__name__", "")
    for kind in ("voc", "coco", "visdrone"):
        if name.endswith("_" + kind):
            return kind
    return ""


def detect_batch(self, batch_img, batch_img_shape):
    """One batch of Evaluator.evaluate (eval/evaluator.py:50-60) -> list of (K,6) float32 numpy arrays, one per
    image, identical to what the reference's per-image `tools.torch_nms(...).cpu().numpy()` produces on our decode.

    Fast route (model prepared by fuse_eval_concat, CUDA): the raw heads go straight into the fused decode + recover
    + threshold + NMS kernel - ONE launch for the batch, nothing of size B x N written - and the rows come back in ONE
    device->host copy.  Otherwise: predict + recover as the reference does, then one batched NMS launch."""
    from . import fused
    model = self.model
    inner = model.module if isinstance(model, nn.DataParallel) and len(model.device_ids) <= 1 else model
    kind = _dataset_kind(self._recover_bboxes)
    if getattr(type(inner), '_pq_fused_forward', False) and batch_img.is_cuda and kind and not inner.training:
        hints = self.__dict__.setdefault('_pq_hints', fused.StrategyHints())
        size = tuple(float(v) for v in self._input_size)
        shapes = batch_img_shape.to(batch_img.device)
        with torch.no_grad():
            yolos, outs = type(inner)._pq_collect(inner, batch_img)
            convs = [getattr(o, '_pq_pending_conv', None) for o in outs]
            strides, C = [l.opt['stride'] for l in yolos], yolos[0].opt['classes']
            if all(c is not None for c in convs):
                # fuse_head_convs is in place: the head convolutions run inside the detection kernels too
                # (features -> detections, nothing of size B x N is written)
                dets = fused.features_nms(outs, [c.weight for c in convs], [c.bias for c in convs], strides, C, size,
                                          shapes, kind, self._score_threshold, self._iou_threshold, hints=hints)
            else:
                raws = [o if c is None else torch.nn.functional.conv2d(o, c.weight, c.bias) for o, c in zip(outs, convs)]
                dets = fused.decode_nms(raws, strides, C, size, shapes, kind, self._score_threshold,
                                        self._iou_threshold, hints=hints)
        return dets.to_numpy_list()
    batch_pred_bbox = self.predict(batch_img)
    device = batch_pred_bbox.device
    input_size = torch.FloatTensor(self._input_size).to(device)
    rec = self._recover_bboxes(batch_pred_bbox, input_size, batch_img_shape.to(device))
    if rec.is_cuda:
        outs = pq_tools.batched_torch_nms(rec, self._score_threshold, self._iou_threshold)
        return [o.cpu().numpy() for o in outs]
    ref_tools = sys.modules.get("tools") or importlib.import_module("tools")
    return [ref_tools.torch_nms(p, self._score_threshold, self._iou_threshold).cpu().numpy() for p in rec]


def _make_evaluate(tqdm):
    def evaluate(self):
        it = tqdm(self.dataset) if tqdm is not None else self.dataset
        for data in it:
            batch_img, batch_file_name, batch_img_shape, batch_label, batch_diff = data
            rows = detect_batch(self, batch_img, batch_img_shape)
            for file_name, labels, diffs, bboxes in zip(batch_file_name, batch_label, batch_diff, rows):
                # the reference hands add_detections an array of shape (0,) for an empty image
                self.add_detections(file_name, bboxes if len(bboxes) else bboxes.reshape(0))
                self.add_labels(file_name, labels, diffs)
        return self.AP()
    return evaluate
