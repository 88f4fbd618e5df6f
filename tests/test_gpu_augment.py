"""-m gpu: eval pre-processing (pqdet_b200/augment.py -> pqdet_letterbox_normalize) against the reference's
Resize -> Normalize -> ToTensor output (tests/golden/letterbox.npz) and the oracle; uint8 and float32 bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu


def test_letterbox_golden_reference_chain():
    from pqdet_b200 import augment as pa
    g = load_golden("letterbox")
    th, tw = (int(v) for v in g["target_hw"])
    imgs = [g["img%d" % i] for i in range(int(g["n"]))]
    out, u8, geo = pa.letterbox_normalize(imgs, (th, tw), want_uint8=True)
    assert tuple(out.shape) == (len(imgs), 3, th, tw) and out.dtype == torch.float32
    for i in range(len(imgs)):
        assert np.array_equal(u8[i].cpu().numpy(), g["resized%d" % i]), i
        assert np.array_equal(out[i].cpu().numpy(), g["tensor%d" % i]), i
        ratio, du, dl = geo[i]
        assert np.array_equal(pa.resize_bboxes(g["bb%d" % i].copy(), ratio, du, dl), g["rbb%d" % i])
    img1, bb1 = pa.Resize((th, tw))(imgs[0].copy(), g["bb0"].copy())            # the reference's call signature
    assert np.array_equal(img1, g["resized0"]) and np.array_equal(bb1, g["rbb0"])


@pytest.mark.parametrize("T", [320, 512, 608])
def test_letterbox_vs_oracle_voc_sized_batch(T):
    from pqdet_b200 import augment as pa
    rng = np.random.default_rng(T)
    shapes = [(375, 500), (500, 333), (281, 500), (500, 500), (112, 640), (1080, 1920), (7, 9), (T, T)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    out, u8, _ = pa.letterbox_normalize(imgs, T, want_uint8=True)
    only_f, _ = pa.letterbox_normalize(imgs, T)
    assert torch.equal(out, only_f)
    for i, im in enumerate(imgs):
        padded, _ = ao.resize_letterbox(im, (T, T))
        assert np.array_equal(u8[i].cpu().numpy(), padded), shapes[i]
        assert np.array_equal(out[i].cpu().numpy(), ao.normalize_to_chw(padded, pa.VOC_MEAN, pa.VOC_STD)), shapes[i]


def test_letterbox_edge_cases():
    from pqdet_b200 import augment as pa
    out, geo = pa.letterbox_normalize([], 64)
    assert tuple(out.shape) == (0, 3, 64, 64) and geo == []
    with pytest.raises(TypeError):
        pa.letterbox_normalize([np.zeros((4, 4, 3), np.float32)], 64)
    one = np.full((1, 1, 3), 200, np.uint8)                                    # a single pixel blown up to 64x64
    out, u8, _ = pa.letterbox_normalize([one], 64, want_uint8=True)
    assert bool((u8 == 200).all())


def test_visdrone_resize_ratio_pad_golden_and_oracle():
    """eval_augment_visdrone: ResizeRatio(1.25) -> PadNearestDivisor(128, 32) -> Normalize -> ToTensor; every image has
    its own canvas."""
    from pqdet_b200 import augment as pa
    g = load_golden("letterbox")
    imgs = [g["vis_img%d" % i] for i in range(int(g["vis_n"]))]
    outs, u8s, info = pa.resize_ratio_pad_normalize(imgs, want_uint8=True)
    for i in range(len(imgs)):
        assert np.array_equal(u8s[i].cpu().numpy(), g["vis_padded%d" % i]), i
        assert np.array_equal(outs[i].cpu().numpy(), g["vis_tensor%d" % i]), i
        bb = g["vis_bb%d" % i].copy()                                  # augment.py:270-272, 295-297
        bb[:, [0, 2]] *= info[i][1]; bb[:, [1, 3]] *= info[i][0]
        bb[:, [0, 2]] += info[i][3]; bb[:, [1, 3]] += info[i][2]
        assert np.array_equal(bb, g["vis_rbb%d" % i])
    rng = np.random.default_rng(9)
    big = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((540, 960), (540, 960), (765, 1360), (101, 77))]
    outs, u8s, _ = pa.resize_ratio_pad_normalize(big, want_uint8=True)
    for k, im in enumerate(big):
        padded, _ = ao.resize_ratio_pad(im)
        assert np.array_equal(u8s[k].cpu().numpy(), padded), im.shape
        assert np.array_equal(outs[k].cpu().numpy(), ao.normalize_to_chw(padded, pa.VOC_MEAN, pa.VOC_STD))
