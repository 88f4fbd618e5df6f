"""Randomised sweep of the materialising decode and of the head conv + decode kernels on a GPU box (not collected by
pytest; about a minute):

    python tests/stress_gpu_decode.py [n_cases] [seed]

Every case draws shapes on both sides of the fast-path conditions (H*W a multiple of 128 or not, class counts from 0 to
80, input channels that do or do not fit the resident-weights plan, batches with more or fewer tiles than SMs) and checks
the persistent kernels (TMA decode pipeline, warp-specialised tcgen05 head conv) bit for bit against the general kernels
(PQDET_DECODE_GENERAL / PQDET_HEADCONV_GENERAL), the decoded rows against the numpy oracle within the 1e-5 contract, and
the head conv's raw output against an fp64 convolution within TF32 precision."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pqdet_oracle as po  # noqa: E402
from pqdet_b200 import _ops  # noqa: E402


def bits(t):
    return t.contiguous().view(torch.int32)


def decode_case(rng):
    C = int(rng.choice([0, 1, 2, 3, 5, 10, 20, 40, 80]))
    A = 3
    ch = 5 + C
    n_levels = int(rng.integers(1, 4))
    aligned = rng.random() < 0.7
    levels = []
    for _ in range(n_levels):
        if aligned:
            H, W = [(8, 16), (16, 16), (16, 32), (32, 32), (64, 64), (4, 32), (128, 8)][int(rng.integers(0, 7))]
        else:
            H, W = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        levels.append((H, W))
    B = int(rng.choice([1, 2, 3, 7, 33]))
    strides = [float(rng.choice([8, 16, 32])) for _ in levels]
    raws = [torch.randn((B, A * ch, H, W), device="cuda") * 2.0 for H, W in levels]
    if rng.random() < 0.3:
        raws[0][0, :, 0, 0] = -100.0
    got = _ops.decode_levels(raws, C, strides)
    os.environ["PQDET_DECODE_GENERAL"] = "1"
    ref = _ops.decode_levels(raws, C, strides)
    single = [_ops.decode_fwd(r, C, s).reshape(B, -1, ch) for r, s in zip(raws, strides)]
    del os.environ["PQDET_DECODE_GENERAL"]
    torch.cuda.synchronize()
    assert torch.equal(bits(got), bits(ref)), ("decode_levels", C, levels, B)
    assert torch.equal(bits(torch.cat(single, 1)), bits(got)), ("decode_fwd", C, levels, B)
    # oracle (numpy fp32 restatement of Decode.forward), first level, first image
    want = po.decode(raws[0][:1].cpu().numpy(), C, strides[0]).reshape(1, -1, ch)
    n0 = levels[0][0] * levels[0][1] * A
    g0 = got[:1, :n0].cpu().numpy()
    fin = np.isfinite(want)
    scale = max(float(levels[0][0]), float(levels[0][1])) * strides[0]
    assert np.all(np.abs(g0[fin] - want[fin]) <= 1e-5 * np.maximum(np.abs(want[fin]), scale)), ("oracle", C, levels)
    return "decode C=%d levels=%s B=%d %s" % (C, levels, B, "aligned" if aligned else "ragged")


def headconv_case(rng):
    C = int(rng.choice([0, 1, 3, 5, 10, 20, 27, 40, 80]))
    ACH = 3 * (5 + C)
    Cin = int(rng.choice([8, 16, 24, 40, 80, 96, 176, 352, 30, 7, 512]))
    if rng.random() < 0.7:
        H, W = [(8, 16), (16, 16), (16, 32), (32, 32), (64, 64), (4, 32), (38, 38), (76, 76), (19, 19), (20, 26)][int(rng.integers(0, 10))]
    else:
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    B = int(rng.choice([1, 2, 5, 40]))
    if B * Cin * H * W > 40e6:
        B = 1
    stride = float(rng.choice([8, 16, 32]))
    x = torch.randn((B, Cin, H, W), device="cuda")
    w = torch.randn((ACH, Cin, 1, 1), device="cuda") * 0.05
    bias = torch.randn((ACH,), device="cuda") * 0.1 if rng.random() < 0.8 else None
    dec, raw = _ops.head_conv_decode(x, w, bias, C, stride, want_raw=True)
    dec_only = _ops.head_conv_decode(x, w, bias, C, stride)
    os.environ["PQDET_HEADCONV_GENERAL"] = "1"
    dec_g, raw_g = _ops.head_conv_decode(x, w, bias, C, stride, want_raw=True)
    del os.environ["PQDET_HEADCONV_GENERAL"]
    torch.cuda.synchronize()
    assert torch.equal(bits(raw), bits(raw_g)) and torch.equal(bits(dec), bits(dec_g)), ("headconv", C, Cin, H, W, B)
    assert torch.equal(bits(dec_only), bits(dec_g)), ("headconv dec only", C, Cin, H, W, B)
    ref = torch.einsum("bchw,oc->bohw", x.double(), w.view(ACH, Cin).double())
    bound = 2.0 ** -9 * torch.einsum("bchw,oc->bohw", x.double().abs(), w.view(ACH, Cin).double().abs()) + 1e-5
    if bias is not None:
        ref = ref + bias.double().view(1, -1, 1, 1)
    assert bool(((raw.double() - ref).abs() <= bound).all()), ("tf32 bound", C, Cin, H, W, B)
    assert torch.equal(bits(dec), bits(_ops.decode_fwd(raw, C, stride))), ("decode of raw", C, Cin, H, W, B)
    return "headconv C=%d Cin=%d %dx%d B=%d" % (C, Cin, H, W, B)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    for i in range(n):
        what = decode_case(rng) if i % 2 == 0 else headconv_case(rng)
        print("[%d/%d] ok  %s" % (i + 1, n, what), flush=True)
    print("all %d cases passed" % n)


if __name__ == "__main__":
    main()
