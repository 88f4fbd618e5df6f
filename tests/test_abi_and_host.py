"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header declares,
the product never imports the oracle, CPU tensors are rejected (no fallback), and the host-side helpers
(shard ranges, synthetic generator, error mapping) behave."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from pqdet_b200 import build, _lib
    build.build()                      # nvcc cross-compiles without a GPU
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "pqdet_b200.h")).read()
    declared = set(re.findall(r"\b(pqdet_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"pqdet_heads_t"}
    from pqdet_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.pqdet_version() == 111
    assert lib.pqdet_strerror(0) == b"ok" and b"unsupported" in lib.pqdet_strerror(-3)


def test_argument_validation_happens_before_any_cuda_call(lib):
    # NULL pointers / bad sizes must come back as error codes, never crash (no GPU needed)
    assert lib.pqdet_decode_fwd(None, None, 1, 3, 20, 4, 4, ctypes.c_float(8), 48, 0, 0, None) == -1
    assert lib.pqdet_recover(None, None, 1, 10, 20, 7, ctypes.c_float(1), ctypes.c_float(1), None, 0, 0, None) == -1
    assert lib.pqdet_nms_general_workspace(4, 16128, 20, 1 << 16, 1) > 0
    assert lib.pqdet_loss_workspace(16, 3, 64, 64) >= 16 * 128 * 3 * 8
    H = (ctypes.c_int * 3)(64, 32, 16)
    assert lib.pqdet_assign_workspace(2, H, H) >= 2 * (64 * 64 + 32 * 32 + 16 * 16) * 3 * 4
    assert lib.pqdet_loss_levels_workspace(3, 16, 3, H, H) > 0
    # the section 8f entry points: same contract
    assert lib.pqdet_head_conv_decode(None, None, None, None, None, 1, 8, 16, 16, 3, 20, ctypes.c_float(8), 768, 0, 0,
                                      None) == -1
    assert lib.pqdet_head_conv_decode(None, None, None, None, None, 0, 8, 16, 16, 3, 20, ctypes.c_float(8), 768, 0, 0,
                                      None) == 0                      # empty batch: nothing to do, no pointer needed
    assert lib.pqdet_head_conv_decode_levels(0, None, None, None, None, None, None, None, None, 1, 3, 20, 0, None) == -1
    assert lib.pqdet_head_conv_decode_levels(5, None, None, None, None, None, None, None, None, 1, 3, 20, 0, None) == -1
    assert lib.pqdet_decode_levels(0, None, None, None, None, None, 1, 3, 20, 0, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pqdet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "/root/reference" not in txt, f
                assert "torchvision" not in txt or f in ("config.py", "tools.py", "fused.py", "nms.cu", "pq_math.cuh"), f
    # torchvision may be *mentioned* in comments of the files above but never imported
    for f in ("config.py", "tools.py", "fused.py"):
        assert not re.search(r"^\s*(from|import)\s+torchvision", open(os.path.join(pkg, f)).read(), re.M)


def test_cpu_tensors_raise_everywhere(lib):
    from pqdet_b200 import base_sample, fused, tools
    from pqdet_b200._lib import PqdetError
    from pqdet_b200.parser import Decode, YOLOLayer
    x = torch.zeros(1, 75, 4, 4)
    with pytest.raises(PqdetError):
        Decode(20, 8)(x)
    with pytest.raises(PqdetError):
        tools.torch_nms(torch.zeros(8, 24), 0.1, 0.45)
    with pytest.raises(PqdetError):
        base_sample.recover_bboxes_prediction_voc(torch.zeros(1, 8, 25), (512, 512), torch.tensor([512., 512.]))
    with pytest.raises(PqdetError):
        fused.decode_nms([x], [8], 20, (32, 32), torch.tensor([32., 32.]))
    with pytest.raises(PqdetError):
        opt = dict(classes=20, stride=8, bbox_loss="l1", ignore_thresh=0.5, l1_loss_gain=0.05)
        YOLOLayer(opt)(x, (torch.zeros(1, 4, 4, 3, 26), torch.zeros(1, 1, 4)))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pqdet_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib._build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PqdetError, match="no CPU fallback"):
        _lib.load()


def test_shard_range_partitions_the_batch():
    from pqdet_b200.dist import shard_range
    for total in (1024, 1000, 7):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synth_generator_shapes_and_determinism():
    from pqdet_b200 import synth
    a = synth.make_heads(2, 20, 512, "sparse", seed=3)
    b = synth.make_heads(2, 20, 512, "sparse", seed=3)
    assert [tuple(t.shape) for t in a] == [(2, 75, 16, 16), (2, 75, 32, 32), (2, 75, 64, 64)]
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    c = synth.make_heads(1, 10, 608, "dense", seed=1)
    assert [tuple(t.shape) for t in c] == [(1, 45, 19, 19), (1, 45, 38, 38), (1, 45, 76, 76)]
    gt = synth.make_gt(3, 20, 512, 1, 12, seed=0)
    assert all(g.shape[1] == 6 and 1 <= len(g) <= 12 for g in gt)


def test_build_center_grid_matches_reference_orientation():
    from pqdet_b200.parser import build_center_grid
    g = build_center_grid(2, 3)
    assert tuple(g.shape) == (2, 3, 1, 2)
    assert g[..., 0, 0].tolist() == [[0.5, 1.5, 2.5], [0.5, 1.5, 2.5]]       # x along W
    assert g[..., 0, 1].tolist() == [[0.5, 0.5, 0.5], [1.5, 1.5, 1.5]]       # y along H


def test_route_hints_record_the_capacity_a_workload_needs():
    """fused._next_route: what a complete run's candidate counts say about the next call (the hint stored in a
    caller-held StrategyHints): compact lists while every image fits them, the large lists while every image fits
    those and the mean stays at half of them, the bucketed general path beyond."""
    from pqdet_b200 import _lib, fused
    (h_c, m_c), (h_l, m_l) = _lib.CAPACITY_LIMITS["compact"], _lib.CAPACITY_LIMITS["large"]
    assert (h_c, m_c, h_l, m_l) == (512, 1280, 1024, 2048)
    assert fused._next_route(0, 0.0) == "compact"
    assert fused._next_route(m_c, 600.0) == "compact"
    assert fused._next_route(m_c + 1, 600.0) == "large"
    assert fused._next_route(m_l, m_l / 2) == "large"
    assert fused._next_route(m_l, m_l / 2 + 1) == "general"
    assert fused._next_route(m_l + 1, 10.0) == "general"
    hints = fused.StrategyHints()
    assert isinstance(hints, dict) and len(hints) == 0


def test_peer_wait_validates_its_arguments(lib):
    """pqdet_peer_wait (the receiving side of the arrival counters): argument checks before any CUDA call."""
    import ctypes
    assert lib.pqdet_peer_wait(None, 2, 0, None, 0, None) == -1            # PQDET_ERR_INVALID_ARG
    buf = (ctypes.c_uint32 * 8)()
    assert lib.pqdet_peer_wait(ctypes.cast(buf, ctypes.c_void_p), 0, 0, None, 0, None) == -1
    assert lib.pqdet_peer_wait(ctypes.cast(buf, ctypes.c_void_p), 9, 0, None, 0, None) == -1
