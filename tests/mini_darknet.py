"""TEST INFRASTRUCTURE: a few-layer stand-in for the reference's cfg-built DetectionModel, for the GPU box (where
/root/reference does not exist).  It restates the *protocol* install.py's hooks rely on - model/interpreter.py:22-85:
`module_list` of layers carrying `_type` (and `_from` / `_layers` / `_stride`), a base class whose forward is the layer
loop returning the [yolo] outputs, a DetectionModel subclass that concatenates (eval) or sums (train) them - around
pqdet_b200.parser.YOLOLayer.  tests/test_install_hooks.py checks the same hooks against the real reference classes on
CPU; this model lets tests/test_gpu_hooks.py run them end to end on CUDA."""
import torch
from torch import nn

from pqdet_b200.interpreter import _TARGET_MAP
from pqdet_b200.parser import YOLOLayer


def _conv(cin, cout, k, stride=1, act=True, name='convolutional'):
    seq = nn.Sequential()
    seq.add_module('conv', nn.Conv2d(cin, cout, k, stride, k // 2, bias=not act))
    if act:
        seq.add_module('bn', nn.BatchNorm2d(cout))
        seq.add_module('act', nn.LeakyReLU(0.1))
    seq._type = name
    return seq


class _Route(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self._type, self._layers = 'route', layers

    def forward(self, xs):
        return xs[0] if len(xs) == 1 else torch.cat(xs, dim=1)


class LoopModel(nn.Module):
    """The layer loop (the reference's AnyModel.forward)."""

    def __init__(self, num_classes=4, bbox_loss='giou', width=16):
        super().__init__()
        C, ch = num_classes, 3 * (5 + num_classes)
        w = width

        def yolo(stride):
            l = YOLOLayer(dict(classes=C, stride=stride, bbox_loss=bbox_loss, ignore_thresh=0.5, l1_loss_gain=0.05))
            l._type, l._stride = 'yolo', stride
            return l
        self.module_list = nn.ModuleList([
            _conv(3, w, 3, 2), _conv(w, w, 3, 2), _conv(w, 2 * w, 3, 2),          # 0-2: stride 8 features
            _conv(2 * w, 2 * w, 3, 2),                                             # 3: stride 16
            _conv(2 * w, 4 * w, 3, 2),                                             # 4: stride 32
            _conv(4 * w, ch, 1, act=False), yolo(32),                              # 5, 6
            _Route([3]), _conv(2 * w, ch, 1, act=False), yolo(16),                 # 7, 8, 9
            _Route([2]), _conv(2 * w, ch, 1, act=False), yolo(8),                  # 10, 11, 12
        ])

    def is_output(self, i, layer):
        return False

    def forward(self, x, target=None):
        cache, outputs = [], []
        for i, layer in enumerate(self.module_list):
            t = layer._type
            if t == 'convolutional':
                x = layer(x)
            elif t == 'route':
                x = layer([cache[j] for j in layer._layers])
            elif t == 'yolo':
                x = layer(x, _TARGET_MAP[layer._stride](target))
            else:
                raise ValueError(t)
            if self.is_output(i, layer):
                outputs.append(x)
            cache.append(x)
        if len(outputs) == 0:
            return cache[-1]
        return outputs[0] if len(outputs) == 1 else outputs


class MiniDetectionModel(LoopModel):
    """model/interpreter.py:67-85."""

    def is_output(self, i, layer):
        return layer._type == 'yolo'

    def forward(self, x, target=None):
        outputs = super().forward(x, target)
        if target is None:
            return torch.cat([o.view((o.shape[0], -1, o.shape[-1])) for o in outputs], dim=1)
        losses = list(map(sum, zip(*outputs)))
        return {'loss': losses[0], 'giou_loss': losses[1], 'conf_loss': losses[2], 'class_loss': losses[3],
                'loss_per_branch': [sum(l[1:]) for l in outputs]}
