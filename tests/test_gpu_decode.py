"""-m gpu: decode / recover kernels vs the oracle and the golden fixtures (through the C ABI)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_close
from gpu_util import cuda
from oracle import pqdet_oracle as po
from oracle import loss_ref

pytestmark = pytest.mark.gpu


def test_decode_golden():
    from pqdet_b200.parser import Decode
    g = load_golden("decode")
    C = int(g["num_classes"])
    for s in (32, 16, 8):
        out = Decode(C, s)(torch.from_numpy(g["raw_s%d" % s]).cuda()).cpu().numpy()
        want = g["out_s%d" % s]
        assert out.shape == want.shape
        # tolerance: 1e-5 relative (north_star); boxes relative to the input extent (fp32 ulp of a coordinate)
        assert rel_close(out[..., :4], want[..., :4], 1e-5, scale=float(s * max(out.shape[1:3])))
        assert rel_close(out[..., 4:], want[..., 4:], 1e-5, scale=1e-30)
    z = Decode(20, 8)(torch.zeros(1, 75, 2, 3).cuda())[0, 1, 2, 0].cpu().numpy()
    assert list(z[:6]) == [12.0, 4.0, 28.0, 20.0, 0.5, 0.5]


@pytest.mark.parametrize("B,C,H,W,s", [(2, 20, 16, 16, 32), (3, 10, 19, 19, 32), (1, 80, 38, 38, 16),
                                        (2, 1, 5, 7, 8), (1, 20, 64, 64, 8), (4, 3, 33, 31, 8)])
def test_decode_vs_oracle(B, C, H, W, s):
    from pqdet_b200.parser import Decode
    g = torch.Generator().manual_seed(B * 1000 + C)
    raw = torch.randn((B, 3 * (5 + C), H, W), generator=g) * 1.5
    out = Decode(C, s)(raw.cuda()).cpu().numpy()
    want = po.decode(raw.numpy(), C, s)
    assert out.shape == want.shape == (B, H, W, 3, 5 + C)
    assert rel_close(out[..., :4], want[..., :4], 1e-5, scale=float(s * max(H, W)))
    assert rel_close(out[..., 4:], want[..., 4:], 1e-5, scale=1e-30)


def test_detection_head_eval_concat_and_known_shape():
    from pqdet_b200.interpreter import DetectionHead
    from pqdet_b200 import synth
    C, size, B = 20, 512, 2
    heads = synth.make_heads(B, C, size, "sparse", seed=3)
    opts = [dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in synth.FPN_STRIDES]
    out = DetectionHead(opts)([h.cuda() for h in heads])
    assert tuple(out.shape) == (B, 16128, 25)                 # export/onnx_exporter.py:373
    want = po.detect([h.numpy() for h in heads], C, synth.FPN_STRIDES)
    got = out.cpu().numpy()
    assert rel_close(got[..., :4], want[..., :4], 1e-5, scale=float(size))
    assert rel_close(got[..., 4:], want[..., 4:], 1e-5, scale=1e-30)


def test_decode_backward_vs_autograd():
    from pqdet_b200.parser import Decode
    C, s = 4, 16
    g = torch.Generator().manual_seed(5)
    raw = torch.randn((2, 3 * (5 + C), 6, 5), generator=g)
    w = torch.randn((2, 6, 5, 3, 5 + C), generator=g)
    r1 = raw.clone().requires_grad_(True)
    (loss_ref.decode_t(r1, C, s) * w).sum().backward()
    r2 = raw.clone().cuda().requires_grad_(True)
    (Decode(C, s)(r2) * w.cuda()).sum().backward()
    gw = r1.grad.numpy()
    assert rel_close(r2.grad.cpu().numpy(), gw, 1e-5, scale=float(np.abs(gw).max()))


def test_recover_golden_and_oracle_bit_exact():
    from pqdet_b200 import base_sample
    g = load_golden("recover")
    pred = torch.from_numpy(g["pred"]).cuda()
    inp = tuple(g["input_size"].tolist())
    orig = torch.from_numpy(g["orig"]).cuda()
    for kind in ("voc", "coco", "visdrone"):
        out = base_sample.RECOVER_BBOXES_REGISTER[kind](pred.clone(), inp, orig).cpu().numpy()
        assert np.array_equal(out, g["out_" + kind]), kind
    out = base_sample.recover_bboxes_prediction_voc(pred.clone(), torch.tensor(inp), orig[0]).cpu().numpy()
    assert np.array_equal(out, g["out_voc_1d"])
    out = base_sample.recover_bboxes_prediction_voc(pred.clone(), tuple(g["input_size2"].tolist()), orig).cpu().numpy()
    assert np.array_equal(out, g["out_voc_rect"])
    # random, larger, all kinds
    gen = torch.Generator().manual_seed(9)
    B, N, C = 5, 3000, 10
    p = torch.rand((B, N, 5 + C), generator=gen)
    p[..., :4] = p[..., :4] * 700 - 50
    o = torch.tensor([[480., 640.], [1080., 1920.], [375., 500.], [500., 375.], [608., 608.]])
    for kind in ("voc", "coco", "visdrone"):
        got = base_sample.RECOVER_BBOXES_REGISTER[kind](p.cuda(), (608., 608.), o.cuda()).cpu().numpy()
        assert np.array_equal(got, po.recover(p.numpy(), (608., 608.), o.numpy(), kind)), kind


def test_cpu_tensor_is_rejected():
    from pqdet_b200.parser import Decode
    from pqdet_b200._lib import PqdetError
    with pytest.raises(PqdetError):
        Decode(20, 8)(torch.zeros(1, 75, 2, 2))


def test_empty_batch_through_the_materialising_route():
    """B = 0 tensors have no storage (data_ptr 0): decode, the eval concat and recover return empty results."""
    from pqdet_b200 import base_sample
    from pqdet_b200.interpreter import DetectionHead
    from pqdet_b200.parser import Decode
    C = 20
    assert tuple(Decode(C, 8)(torch.zeros((0, 75, 64, 64), device="cuda")).shape) == (0, 64, 64, 3, 25)
    head = DetectionHead([dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in (32, 16, 8)])
    pred = head([torch.zeros((0, 75, 512 // s, 512 // s), device="cuda") for s in (32, 16, 8)])
    assert tuple(pred.shape) == (0, 16128, 25)
    rec = base_sample.recover_bboxes_prediction_voc(pred, (512, 512), torch.zeros((0, 2), device="cuda"))
    assert tuple(rec.shape) == (0, 16128, 24)


@pytest.mark.parametrize("B,Cin,H,W,C,stride", [(2, 80, 64, 64, 20, 8), (3, 352, 16, 16, 20, 32), (2, 176, 19, 19, 10, 16),
                                                  (1, 80, 38, 38, 80, 8), (2, 30, 7, 9, 1, 16),
                                                  (9, 176, 64, 64, 20, 8), (5, 40, 32, 32, 10, 16), (2, 96, 16, 24, 3, 16)])
def test_head_conv_decode_tensor_cores(B, Cin, H, W, C, stride):
    """1x1 head convolution + Decode on tcgen05 (TF32 products, fp32 accumulation in TMEM).
    raw vs an fp64 convolution within TF32 precision (operands truncated to 10 mantissa bits: up to 2^-10 each); the decoded
    output is bit-identical to Decode applied to the kernel's own raw output (same epilogue arithmetic)."""
    from pqdet_b200 import _ops
    g = torch.Generator(device="cuda").manual_seed(Cin + H)
    A = 3
    ACH = A * (5 + C)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((ACH, Cin, 1, 1), device="cuda", generator=g) * 0.05
    bias = torch.randn((ACH,), device="cuda", generator=g) * 0.1
    dec, raw = _ops.head_conv_decode(x, w, bias, C, stride, want_raw=True)
    # oracle: fp64 convolution; |error| <= 2 * 2^-10 * sum_c |x_c * w_c| (the tensor core truncates both operands to
    # TF32; measured worst ratio 1.4e-3) + fp32 accumulation noise
    from oracle import pqdet_oracle as po
    xn, wn = x.cpu().numpy(), w.cpu().numpy()
    ref = po.head_conv(xn, wn, bias.cpu().numpy())
    assert np.all(np.abs(raw.cpu().numpy() - ref) <= po.head_conv_error_bound(xn, wn, 2.0 ** -10))
    assert torch.equal(dec, _ops.decode_fwd(raw, C, stride))
    want = po.decode(raw.cpu().numpy(), C, stride)                       # oracle Decode of the kernel's own raw head
    got = dec.cpu().numpy()
    assert np.all(np.abs(got - want) <= 1e-5 * np.maximum(np.abs(want), max(H, W) * stride))
    no_bias = _ops.head_conv_decode(x, w, None, C, stride, want_raw=True)[1]
    assert torch.allclose(no_bias + bias.view(1, -1, 1, 1), raw, rtol=0, atol=1e-5)


@pytest.mark.parametrize("B,Cin,H,W,C", [(3, 352, 16, 16, 20), (7, 176, 32, 32, 20), (40, 80, 64, 64, 20), (2, 48, 16, 16, 10),
                                         (150, 80, 16, 16, 40), (3, 80, 16, 16, 80), (5, 8, 32, 16, 1), (2, 1024, 16, 8, 20),
                                         (4, 64, 16, 16, 5), (3, 32, 16, 16, 0),
                                         # 255 channels: anchor-sliced epilogue (weights fit) and CTAs split between the
                                         # anchors (Cin 176 / 352: they do not), partial last tiles, one image, W != H
                                         (1, 176, 38, 38, 80), (2, 80, 24, 40, 80), (3, 176, 20, 28, 80), (2, 352, 12, 20, 80),
                                         (5, 176, 76, 76, 80)])
def test_head_conv_persistent_kernel_equals_general_kernel(B, Cin, H, W, C, monkeypatch):
    """H*W a multiple of 4 takes the persistent warp-specialised kernel (TMA ring, resident weights, two TMEM
    accumulators); PQDET_HEADCONV_GENERAL forces the general kernel.  Same K order of the same tf32 MMAs: identical
    raw heads and decoded rows; more tiles than SMs (40 x 32) exercises the ring / accumulator phases."""
    from pqdet_b200 import _ops
    g = torch.Generator(device="cuda").manual_seed(B + Cin)
    ACH = 3 * (5 + C)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((ACH, Cin, 1, 1), device="cuda", generator=g) * 0.05
    bias = torch.randn((ACH,), device="cuda", generator=g) * 0.1
    dec, raw = _ops.head_conv_decode(x, w, bias, C, 8.0, want_raw=True)
    dec_only = _ops.head_conv_decode(x, w, bias, C, 8.0)
    dec_only = dec_only[0] if isinstance(dec_only, tuple) else dec_only
    monkeypatch.setenv("PQDET_HEADCONV_GENERAL", "1")
    dec_g, raw_g = _ops.head_conv_decode(x, w, bias, C, 8.0, want_raw=True)
    torch.cuda.synchronize()
    assert torch.equal(raw, raw_g)
    assert torch.equal(dec, dec_g) and torch.equal(dec_only, dec_g)


@pytest.mark.parametrize("B,C,size,strides", [(3, 20, 512, (32, 16, 8)), (37, 20, 512, (32, 16, 8)), (5, 10, 512, (32, 16, 8)),
                                              (2, 1, 512, (32, 16, 8)), (2, 80, 512, (32, 16, 8)), (9, 3, 1024, (32, 16, 8)),
                                              (3, 10, 608, (32, 16, 8)), (2, 20, 416, (32, 16, 8)), (2, 20, 320, (32, 16, 8)),
                                              (3, 10, 608, (8, 16, 32)), (2, 20, 416, (8, 16, 32)), (40, 10, 608, (32, 16, 8))])
def test_eval_concat_tma_pipeline_equals_general_kernel(B, C, size, strides, monkeypatch):
    """The eval concat runs as a persistent TMA pipeline (tensor-map loads -> decode -> bulk store) for every level
    whose plane stride a tensor map can address; PQDET_DECODE_GENERAL forces the general kernel.  Identical bits, more
    tiles than SMs included.  C = 80 does not fit the pipeline's shared memory and falls back by itself; 608 / 416 /
    320 inputs have a level the pipeline cannot take (19x19, 13x13: general kernel, same launch sequence), partial
    last tiles, and - behind such a level - row ranges that are not 16-byte aligned (cooperative stores)."""
    from pqdet_b200 import _ops
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C)
    ACH = 3 * (5 + C)
    raws = [torch.randn((B, ACH, size // s, size // s), device="cuda", generator=g) * 1.5 for s in strides]
    raws[0][0, :, 0, 0] = -120.0                                     # the checked reciprocal path
    raws[1][-1, 4::5 + C, 1, 2] = float("nan")
    got = _ops.decode_levels(raws, C, strides)
    got2 = _ops.decode_levels(raws, C, strides)                      # staging tiles / ring phases reused
    monkeypatch.setenv("PQDET_DECODE_GENERAL", "1")
    ref = _ops.decode_levels(raws, C, strides)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert torch.equal(got.view(torch.int32), ref.view(torch.int32)) and torch.equal(got2.view(torch.int32), ref.view(torch.int32))
    rows = [_ops.decode_fwd(r, C, s).reshape(B, -1, 5 + C) for r, s in zip(raws, strides)]
    assert torch.equal(torch.cat(rows, 1).view(torch.int32), got.view(torch.int32))


def test_single_level_decode_into_a_row_range(monkeypatch):
    """pqdet_decode_fwd with out_rows_total / out_row_offset (the C ABI's own way to build the eval concat) through the
    TMA pipeline and through the general kernel."""
    from pqdet_b200 import _ops
    B, C = 3, 20
    g = torch.Generator(device="cuda").manual_seed(11)
    raws = [torch.randn((B, 75, 512 // s, 512 // s), device="cuda", generator=g) for s in (32, 16, 8)]
    N = sum(r.shape[2] * r.shape[3] * 3 for r in raws)
    outs = []
    for general in (False, True):
        if general:
            monkeypatch.setenv("PQDET_DECODE_GENERAL", "1")
        out = torch.full((B, N, 5 + C), -7.0, device="cuda")
        off = 0
        for r, s in zip(raws, (32, 16, 8)):
            _ops.decode_fwd(r, C, s, out=out, rows_total=N, row_offset=off)
            off += r.shape[2] * r.shape[3] * 3
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    monkeypatch.delenv("PQDET_DECODE_GENERAL")
    assert torch.equal(outs[0], _ops.decode_levels(raws, C, (32, 16, 8)))


def test_sigmoid_reciprocal_out_of_range_logits(monkeypatch):
    """The decode kernels interleave several 1/(1+e) chains with the range check of __frcp_rn hoisted out
    (pq_math.cuh rcp_rn_core); logits below -87 (1+e >= 2^126: denormal / zero sigmoid), +-inf-producing values and
    NaN must take the checked path and give exactly what the plain __frcp_rn formulation (general head conv
    kernel) gives."""
    from pqdet_b200 import _ops
    C, Cin, H, W = 20, 16, 16, 16
    ACH = 3 * (5 + C)
    vals = torch.tensor([-104.0, -100.0, -95.0, -88.8, -88.0, -87.4, -87.3, -87.0, -50.0, -1.0, 0.0, 3.0, 50.0, 88.0, 89.0,
                         100.0, float("nan"), float("inf"), float("-inf")], device="cuda")
    bias = vals[torch.arange(ACH, device="cuda") % vals.numel()].contiguous()
    x = torch.randn((2, Cin, H, W), device="cuda")
    w = torch.zeros((ACH, Cin, 1, 1), device="cuda")
    dec_ws, raw_ws = _ops.head_conv_decode(x, w, bias, C, 8.0, want_raw=True)
    assert torch.equal(raw_ws.nan_to_num(7.0), bias.view(1, -1, 1, 1).expand_as(raw_ws).nan_to_num(7.0))
    dec_mat = _ops.decode_fwd(raw_ws, C, 8.0)                        # materialising decode kernel, 4 chains interleaved
    monkeypatch.setenv("PQDET_HEADCONV_GENERAL", "1")
    dec_gen = _ops.head_conv_decode(x, w, bias, C, 8.0)               # plain __frcp_rn per element
    torch.cuda.synchronize()
    for got in (dec_ws, dec_mat):
        assert torch.equal(got.view(torch.int32), dec_gen.view(torch.int32))     # bit patterns, NaN included
    sig = dec_gen.view(2, H, W, 3, 5 + C)[0, 0, 0].reshape(-1)
    ref = 1.0 / (1.0 + torch.exp(-bias.double()))
    keep = (torch.arange(ACH, device="cuda") % (5 + C)) >= 4
    ok = torch.isfinite(ref) & keep
    assert torch.allclose(sig[ok].double(), ref[ok], rtol=1e-5, atol=1e-37)


@pytest.mark.parametrize("B,C,size,cins", [(2, 20, 512, (352, 176, 80)), (37, 20, 512, (352, 176, 80)), (3, 10, 512, (64, 32, 16)),
                                           (1, 20, 512, (352, 176, 80)), (5, 20, 1024, (48, 24)), (100, 20, 512, (80, 40, 16))])
def test_head_conv_all_levels_in_one_launch(B, C, size, cins):
    """pqdet_head_conv_decode_levels: the SMs are split between the levels, every CTA keeps its level's weights
    resident; rows identical to one pqdet_head_conv_decode launch per level."""
    from pqdet_b200 import _ops
    g = torch.Generator(device="cuda").manual_seed(B + C)
    ch = 5 + C
    strides = (32, 16, 8)[:len(cins)]
    feats = [torch.randn((B, c, size // s, size // s), device="cuda", generator=g) for c, s in zip(cins, strides)]
    ws = [torch.randn((3 * ch, c, 1, 1), device="cuda", generator=g) * 0.04 for c in cins]
    bs = [torch.randn((3 * ch,), device="cuda", generator=g) * 0.1 for _ in cins]
    bs[-1] = None
    N = sum(f.shape[2] * f.shape[3] * 3 for f in feats)
    out = torch.full((B, N, ch), -3.0, device="cuda")
    assert _ops.head_conv_decode_levels(feats, ws, bs, C, strides, out)
    ref = torch.empty_like(out)
    off = 0
    for f, w, b, s in zip(feats, ws, bs, strides):
        _ops.head_conv_decode(f, w, b, C, s, out=ref, rows_total=N, row_offset=off)
        off += f.shape[2] * f.shape[3] * 3
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    # a level the persistent kernel cannot take (19x19: plane stride not a multiple of 16 bytes) makes the call
    # decline instead of computing something else; a partial-tile level (8x8 = 64 cells) is computed, identically to
    # the general kernel
    odd = [torch.randn((B, 16, 19, 19), device="cuda")]
    assert not _ops.head_conv_decode_levels(odd, [torch.randn((3 * ch, 16), device="cuda")], [None], C, (32,),
                                            torch.empty((B, 19 * 19 * 3, ch), device="cuda"))
    small, wsm = [torch.randn((B, 16, 8, 8), device="cuda")], [torch.randn((3 * ch, 16), device="cuda")]
    o1 = torch.empty((B, 192, ch), device="cuda")
    if _ops.head_conv_decode_levels(small, wsm, [None], C, (32,), o1):
        import os
        os.environ["PQDET_HEADCONV_GENERAL"] = "1"
        try:
            o2 = _ops.head_conv_decode(small[0], wsm[0], None, C, 32.0).reshape(B, 192, ch)
        finally:
            del os.environ["PQDET_HEADCONV_GENERAL"]
        assert torch.equal(o1, o2)


def test_fuse_head_convs_hook_on_a_reference_shaped_model():
    """install.fuse_head_convs on a model laid out like the reference's DetectionModel (module_list of blocks with
    `_type`, the 1x1 linear head conv in front of every [yolo], model/interpreter.py:40-63): eval output = Decode of the
    convolution within TF32 precision, identical to pqdet_head_conv_decode; training mode, eval with a target, a head
    conv whose output a route reads, and CPU tensors still run the convolution itself."""
    from torch import nn
    from pqdet_b200 import _ops, install, parser as pqp
    C, ch = 20, 25

    class Mock(nn.Module):
        def __init__(self, routed):
            super().__init__()
            def block(mod, t, **kw):
                for k, v in dict(_type=t, **kw).items():
                    setattr(mod, k, v)
                return mod
            trunk = nn.Sequential()
            trunk.add_module('conv', nn.Conv2d(3, 32, 3, padding=1))
            trunk.add_module('act', nn.ReLU())
            head = nn.Sequential()
            head.add_module('conv', nn.Conv2d(32, 3 * ch, 1))
            yolo = pqp.YOLOLayer(dict(classes=C, stride=8, bbox_loss='l1', ignore_thresh=0.5, l1_loss_gain=0.05))
            mods = [block(trunk, 'convolutional'), block(head, 'convolutional'), block(yolo, 'yolo')]
            if routed:
                mods.append(block(nn.Identity(), 'route', _layers=[-2]))
            self.module_list = nn.ModuleList(mods)

        def forward(self, x, target=None):
            cache = []
            out = None
            for layer in self.module_list:
                if layer._type == 'yolo':
                    x = layer(x, target)
                    out = x
                elif layer._type == 'route':
                    x = cache[len(cache) + layer._layers[0]]
                else:
                    x = layer(x)
                cache.append(x)
            return out

    torch.manual_seed(5)
    m = Mock(False).cuda().eval()
    img = torch.randn(2, 3, 32, 64, device="cuda")
    label = torch.zeros((2, 32, 64, 3, 6 + C), device="cuda")            # the mock trunk keeps the resolution
    label[..., 5] = 1.0                                                  # mix weight; no responsible cell
    target = (label, torch.zeros((2, 1, 4), device="cuda"))
    with torch.no_grad():
        want = m(img)                                                    # conv (cuDNN, TF32 allowed) + our Decode
        want_loss = m(img, target)
        assert install.fuse_head_convs(m) == 1
        got = m(img)
        feats = m.module_list[0](img)
        direct = _ops.head_conv_decode(feats, m.module_list[1].conv.weight, m.module_list[1].conv.bias, C, 8)
        assert got.shape == want.shape and torch.equal(got, direct)
        assert torch.allclose(got[..., 4:], want[..., 4:], rtol=0, atol=5e-3)
        # eval with a target (validation loss): the YOLOLayer applies the convolution itself
        got_loss = m(img, target)
        for a_, b_ in zip(got_loss, want_loss):
            assert torch.allclose(a_, b_, rtol=1e-4, atol=1e-6)
        # nn.DataParallel-style replicas (module __dict__ copies) carry no reference to the original modules
        rep = m.module_list[1]._replicate_for_data_parallel()
        assert type(rep).__name__ == '_FusedHeadConv' and not any(v is m.module_list[2] for v in rep.__dict__.values())
        m.train()
        assert m.module_list[1](feats).shape[1] == 3 * ch               # training mode: the block convolves
        m.eval()
    routed = Mock(True).cuda().eval()
    assert install.fuse_head_convs(routed) == 0                         # someone else reads the raw head: left alone


def test_fuse_eval_concat_hook_on_a_reference_shaped_model():
    """install.fuse_eval_concat on a model laid out like the reference's (AnyModel.forward = the layer loop returning
    the [yolo] outputs, DetectionModel.forward = view + cat, model/interpreter.py:40-76): identical bits to the
    per-level Decode + cat; with fuse_head_convs on top, identical to the head-conv kernel per level; training mode and
    targets still take the reference's forward."""
    from torch import nn
    from pqdet_b200 import _ops, install, parser as pqp
    C, ch = 20, 25

    def block(mod, t, **kw):
        for k, v in dict(_type=t, **kw).items():
            setattr(mod, k, v)
        return mod

    class Loop(nn.Module):                       # AnyModel
        def forward(self, x, target=None):
            cache, outputs = [], []
            for layer in self.module_list:
                if layer._type == 'yolo':
                    x = layer(x, target)
                    outputs.append(x)
                elif layer._type == 'route':
                    x = cache[layer._layers[0]]
                else:
                    x = layer(x)
                cache.append(x)
            return outputs

    class Model(Loop):                           # DetectionModel
        def __init__(self):
            super().__init__()
            mods = []
            def conv(cin, cout, k, act):
                b = nn.Sequential()
                b.add_module('conv', nn.Conv2d(cin, cout, k, padding=k // 2))
                if act:
                    b.add_module('act', nn.ReLU())
                return block(b, 'convolutional')
            mods.append(conv(3, 32, 3, True))                                               # 0 trunk
            mods.append(conv(32, 3 * ch, 1, False))                                         # 1 head
            mods.append(block(pqp.YOLOLayer(dict(classes=C, stride=32, bbox_loss='l1', ignore_thresh=0.5)), 'yolo'))
            mods.append(block(nn.Identity(), 'route', _layers=[0]))                         # 3 back to the trunk
            mods.append(block(nn.UpsamplingNearest2d(scale_factor=2), 'upsample'))          # 4
            mods.append(conv(32, 16, 3, True))                                              # 5
            mods.append(conv(16, 3 * ch, 1, False))                                         # 6 head
            mods.append(block(pqp.YOLOLayer(dict(classes=C, stride=16, bbox_loss='l1', ignore_thresh=0.5)), 'yolo'))
            self.module_list = nn.ModuleList(mods)

        def forward(self, x, target=None):
            outputs = super().forward(x, target)
            if target is None:
                return torch.cat([o.view((o.shape[0], -1, o.shape[-1])) for o in outputs], dim=1)
            return outputs

    torch.manual_seed(9)
    m = Model().cuda().eval()
    img = torch.randn(3, 3, 16, 32, device="cuda")
    with torch.no_grad():
        want = m(img)                                                     # Decode per level + view + cat
        assert install.fuse_eval_concat(m) and install.fuse_eval_concat(m)
        got = m(img)
        assert got.shape == want.shape == (3, (16 * 32 + 32 * 64) * 3, ch) and torch.equal(got, want)
        assert install.fuse_head_convs(m) == 2
        fused = m(img)
        f0 = m.module_list[0](img)
        f1 = m.module_list[5](m.module_list[4](f0))
        rows = [_ops.head_conv_decode(f, m.module_list[i].conv.weight, m.module_list[i].conv.bias, C, s).reshape(3, -1, ch)
                for f, i, s in ((f0, 1, 32), (f1, 6, 16))]
        assert torch.equal(fused, torch.cat(rows, 1))
        assert torch.allclose(fused[..., 4:], want[..., 4:], rtol=0, atol=5e-3)
        assert not any('_pq_passthrough' in l.__dict__ for l in m.module_list)
    m.train()
    assert isinstance(m(img, None), torch.Tensor)                        # training mode: the reference's own forward


def test_forward_from_features_equals_conv_then_decode():
    from pqdet_b200 import _ops
    from pqdet_b200.interpreter import DetectionHead
    C, B, size = 20, 2, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    head = DetectionHead([dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in (32, 16, 8)])
    cins = (352, 176, 80)
    feats = [torch.randn((B, c, size // s, size // s), device="cuda", generator=g) for c, s in zip(cins, (32, 16, 8))]
    ws = [torch.randn((75, c, 1, 1), device="cuda", generator=g) * 0.03 for c in cins]
    bs = [torch.randn((75,), device="cuda", generator=g) * 0.1 for _ in cins]
    pred = head.forward_from_features(feats, ws, bs)
    raws = [_ops.head_conv_decode(f, w, b, C, s, want_raw=True)[1] for f, w, b, s in zip(feats, ws, bs, (32, 16, 8))]
    assert torch.equal(pred, head(raws))                       # same rows as Decode + concat of the raw heads
    conv = [torch.nn.functional.conv2d(f, w, b) for f, w, b in zip(feats, ws, bs)]     # PyTorch's own (TF32) conv
    assert torch.allclose(head(conv)[..., 4:], pred[..., 4:], rtol=0, atol=5e-3)


@pytest.mark.parametrize("C", [1, 2, 3, 7])
def test_recover_small_class_counts(C):
    """The score pass maps element -> row by multiply-high with ceil(2^32 / C); C = 1 needs the special case."""
    from pqdet_b200 import base_sample
    rng = np.random.default_rng(C)
    B, N = 3, 4032
    pred = rng.uniform(0, 200, (B, N, 5 + C)).astype(np.float32)
    pred[..., 4:] = rng.uniform(0, 1, (B, N, 1 + C)).astype(np.float32)
    orig = np.array([[375., 500.], [480., 480.], [200., 333.]], np.float32)
    for kind in ("voc", "visdrone"):
        got = base_sample.RECOVER_BBOXES_REGISTER[kind](cuda(pred), (256, 256), cuda(orig)).cpu().numpy()
        assert np.array_equal(got, po.recover(pred, (256, 256), orig, kind)), (C, kind)


@pytest.mark.parametrize("C,size,B,cins,thr", [(20, 512, 6, (352, 176, 80), 0.1), (80, 512, 3, (64, 48, 32), 0.05),
                                               (3, 256, 5, (40, 24, 16), 0.3)])
def test_features_to_detections_equals_raw_head_route(C, size, B, cins, thr):
    """SURVEY 8f-2, second half: head convolution whose epilogue thresholds + the fused kernel's back end on the hit
    records (pqdet_head_conv_hits -> pqdet_records_nms) against the raw-head route on the SAME convolution's raw
    output (pqdet_head_conv_decode(out_raw) -> pqdet_decode_nms): identical rows and indices, and no launch that
    writes a B x N tensor."""
    from pqdet_b200 import _ops, fused, synth
    torch.manual_seed(C + size)
    strides = synth.FPN_STRIDES
    ch = 3 * (5 + C)
    feats = [torch.randn((B, cin, size // s, size // s), device="cuda") for cin, s in zip(cins, strides)]
    ws = [torch.randn((ch, cin), device="cuda") * (1.5 / cin ** 0.5) for cin in cins]
    bs = [torch.randn((ch,), device="cuda") * 0.3 for _ in cins]
    for b_ in bs:
        b_[4::(5 + C)] -= 3.5                                   # objectness logits: a few percent of the rows pass
    orig = torch.tensor([[375., 500.], [333., 500.], [float(size)] * 2, [500., 281.], [480., 640.], [300., 400.]])[:B].cuda()
    raws = [_ops.head_conv_decode(f, w, b_, C, float(s), want_raw=True, want_decoded=False)
            for f, w, b_, s in zip(feats, ws, bs, strides)]
    want = fused.decode_nms(raws, strides, C, (size, size), orig, "voc", thr, 0.45, return_index=True)
    got = fused.features_nms(feats, ws, bs, strides, C, (size, size), orig, "voc", thr, 0.45, return_index=True)
    assert int(want.counts.sum()) > 0
    assert torch.equal(got.host_meta()[0], want.host_meta()[0]) and torch.equal(got.host_meta()[1], want.host_meta()[1])
    for b in range(B):
        assert torch.equal(got[b], want[b]), b
        assert torch.equal(got.indices(b), want.indices(b)), b
    assert len(got._spill) == len(want._spill)


def test_features_to_detections_falls_back_on_unaligned_levels_and_dense_images():
    from pqdet_b200 import _ops, fused, synth
    torch.manual_seed(5)
    C, size, B = 4, 608, 2                                      # 19x19 / 38x38 / 76x76: outside the persistent kernel
    strides = synth.FPN_STRIDES
    ch = 3 * (5 + C)
    cins = (32, 24, 16)
    feats = [torch.randn((B, cin, size // s, size // s), device="cuda") for cin, s in zip(cins, strides)]
    ws = [torch.randn((ch, cin), device="cuda") * 0.2 for cin in cins]
    bs = [torch.randn((ch,), device="cuda") * 0.3 - 1.0 for _ in cins]
    orig = torch.tensor([float(size), float(size)]).cuda()
    raws = [_ops.head_conv_decode(f, w, b_, C, float(s), want_raw=True, want_decoded=False)
            for f, w, b_, s in zip(feats, ws, bs, strides)]
    want = fused.decode_nms(raws, strides, C, (size, size), orig, "coco", 0.1, 0.45)
    got = fused.features_nms(feats, ws, bs, strides, C, (size, size), orig, "coco", 0.1, 0.45)
    for b in range(B):
        assert torch.equal(got[b], want[b])
    # aligned levels, but every row is a hit: the records overflow and the images are resolved through the general path
    size = 256
    feats = [torch.randn((B, cin, size // s, size // s), device="cuda") for cin, s in zip(cins, strides)]
    bs2 = [b_ + 6.0 for b_ in bs]
    raws = [_ops.head_conv_decode(f, w, b_, C, float(s), want_raw=True, want_decoded=False)
            for f, w, b_, s in zip(feats, ws, bs2, strides)]
    want = fused.decode_nms(raws, strides, C, (size, size), orig, "coco", 0.1, 0.45)
    got = fused.features_nms(feats, ws, bs2, strides, C, (size, size), orig, "coco", 0.1, 0.45)
    assert len(got._spill) == B
    for b in range(B):
        assert torch.equal(got[b], want[b])


@pytest.mark.parametrize("C,cins", [(10, (352, 176, 80)), (80, (64, 48, 32)), (80, (352, 176, 80))])
def test_head_conv_608_levels_persistent_vs_general(C, cins, monkeypatch):
    """608 x 608 inputs (BASELINE configs C / D): 38 x 38 and 76 x 76 levels now take the persistent kernel (partial
    last tile through TMA zero fill, rows of the concatenated prediction not 16-byte aligned -> cooperative stores);
    19 x 19 (plane stride 1444 B) stays on the general kernel.  Both kernels must agree bit for bit, per level and
    through forward_from_features (the unaligned row offsets), and the HITS epilogue of both must give the same
    detections."""
    from pqdet_b200 import _ops, fused, synth
    from pqdet_b200.interpreter import DetectionHead
    torch.manual_seed(C)
    B, size = 3, 608
    strides = synth.FPN_STRIDES
    ch = 3 * (5 + C)
    feats = [torch.randn((B, cin, size // s, size // s), device="cuda") for cin, s in zip(cins, strides)]
    ws = [torch.randn((ch, cin), device="cuda") / cin ** 0.5 for cin in cins]
    bs = [torch.randn((ch,), device="cuda") * 0.3 for _ in cins]
    for b_ in bs:
        b_[4::(5 + C)] -= 3.0
    head = DetectionHead([dict(classes=C, stride=s, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05) for s in strides])
    orig = torch.tensor([[480., 480.], [360., 640.], [608., 608.]]).cuda()
    got_levels = [_ops.head_conv_decode(f, w, b_, C, float(s), want_raw=True) for f, w, b_, s in zip(feats, ws, bs, strides)]
    got_cat = head.forward_from_features(feats, ws, bs)
    got_det = fused.features_nms(feats, ws, bs, strides, C, (size, size), orig, "coco", 0.1, 0.45, return_index=True)
    monkeypatch.setenv("PQDET_HEADCONV_GENERAL", "1")
    ref_levels = [_ops.head_conv_decode(f, w, b_, C, float(s), want_raw=True) for f, w, b_, s in zip(feats, ws, bs, strides)]
    ref_cat = head.forward_from_features(feats, ws, bs)
    ref_det = fused.features_nms(feats, ws, bs, strides, C, (size, size), orig, "coco", 0.1, 0.45, return_index=True)
    monkeypatch.delenv("PQDET_HEADCONV_GENERAL")
    torch.cuda.synchronize()
    for (gd, gr), (rd, rr) in zip(got_levels, ref_levels):
        assert torch.equal(gr.view(torch.int32), rr.view(torch.int32)) and torch.equal(gd.view(torch.int32), rd.view(torch.int32))
    assert torch.equal(got_cat.view(torch.int32), ref_cat.view(torch.int32))
    raws = [r for _, r in got_levels]
    want = fused.decode_nms(raws, strides, C, (size, size), orig, "coco", 0.1, 0.45, return_index=True)
    assert int(want.counts.sum()) > 0
    for b in range(B):
        assert torch.equal(got_det[b], want[b]) and torch.equal(ref_det[b], want[b]), b
        assert torch.equal(got_det.indices(b), want.indices(b))
