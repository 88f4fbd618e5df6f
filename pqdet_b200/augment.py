"""Drop-in for the eval pre-processing of dataset/augment.py:
Resize (:227-259) -> Normalize (:206-215) -> ToTensor (:390-398), the chain eval_augment_voc / eval_augment_coco
build (dataset/voc_sample.py:85-90), for a whole batch of uint8 HWC images in one kernel launch (csrc/augment.cu).

The host only does what Resize.__call__ does in Python before calling OpenCV: the letterbox geometry per image
(resize ratio, rounded size, padding) - with the same Python arithmetic - and packs the raw bytes; the bilinear
resize (OpenCV's 8-bit fixed-point arithmetic, reproduced bit-exactly), the padding, the normalisation and the
HWC -> CHW transpose run on the GPU.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

IMAGE_REC = np.dtype([("off", "<i8"), ("sh", "<i4"), ("sw", "<i4"), ("dh", "<i4"), ("dw", "<i4"), ("du", "<i4"),
                      ("dl", "<i4"), ("scale_y", "<f8"), ("scale_x", "<f8")], align=True)
VOC_MEAN, VOC_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)          # dataset/voc_sample.py:88


def letterbox_geometry(img_hw: Tuple[int, int], target_hw: Tuple[int, int]):
    """dataset/augment.py:236-249 -> (resize_ratio, resize_h, resize_w, du, dl)."""
    img_h, img_w = img_hw
    target_h, target_w = target_hw
    resize_ratio = min(target_w / img_w, target_h / img_h)
    resize_w = round(resize_ratio * img_w)
    resize_h = round(resize_ratio * img_h)
    dl = (target_w - resize_w) // 2
    du = (target_h - resize_h) // 2
    return resize_ratio, resize_h, resize_w, du, dl


def resize_bboxes(bboxes: np.ndarray, resize_ratio: float, du: int, dl: int) -> np.ndarray:
    """dataset/augment.py:256-258 (in place, like the reference)."""
    if len(bboxes) != 0:
        bboxes[:, [0, 2]] = bboxes[:, [0, 2]] * resize_ratio + dl
        bboxes[:, [1, 3]] = bboxes[:, [1, 3]] * resize_ratio + du
    return bboxes


_STAGE = {}          # device -> [pinned uint8 staging tensor, event of the last copy out of it]


def _staging(device, nbytes: int) -> np.ndarray:
    """A reusable page-locked staging buffer (allocating pinned memory per call costs more than the copy).  The
    previous asynchronous copy out of it must have completed before it is overwritten."""
    st = _STAGE.get(device)
    if st is not None and st[1] is not None:
        st[1].synchronize()
    if st is None or st[0].numel() < nbytes:
        st = [torch.empty((max(nbytes, 1 << 20),), dtype=torch.uint8, pin_memory=True), None]
        _STAGE[device] = st
    return st[0]


def ratio_pad_geometry(img_hw: Tuple[int, int], ratio=(1.25, 1.25), divisor: int = 32):
    """dataset/augment.py:266-268 (ResizeRatio) + :281-289 (PadNearestDivisor) -> (resize_h, resize_w, canvas_h,
    canvas_w, du, dl): the eval geometry of eval_augment_visdrone (dataset/visdrone_sample.py:76-82)."""
    from math import ceil
    img_h, img_w = img_hw
    resize_h, resize_w = round(ratio[0] * img_h), round(ratio[1] * img_w)
    canvas_h = int(ceil(resize_h / divisor) * divisor)
    canvas_w = int(ceil(resize_w / divisor) * divisor)
    return resize_h, resize_w, canvas_h, canvas_w, (canvas_h - resize_h) // 2, (canvas_w - resize_w) // 2


def _pack(images: Sequence[np.ndarray], target_hw, stage=None, geometry=None):
    recs = np.zeros((len(images),), IMAGE_REC)
    off = 0
    geo = []
    for i, im in enumerate(images):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise TypeError("images must be uint8 HWC with 3 channels, got %s %s" % (im.dtype, im.shape))
        sh, sw = im.shape[:2]
        ratio, dh, dw, du, dl = geometry[i] if geometry is not None else letterbox_geometry((sh, sw), target_hw)
        if dh < 1 or dw < 1:
            raise ValueError("image %d (%dx%d) vanishes at input size %s" % (i, sh, sw, target_hw))
        # cv::resize: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
        recs[i] = (off, sh, sw, dh, dw, du, dl, 1.0 / (dh / sh), 1.0 / (dw / sw))
        geo.append((ratio, du, dl))
        off += sh * sw * 3
    packed = np.empty((off,), np.uint8) if stage is None else stage(off + recs.nbytes + 8)[:off].numpy()
    for r, im in zip(recs, images):
        packed[r["off"]:r["off"] + im.size] = np.ascontiguousarray(im).reshape(-1)
    return packed, recs, geo


def letterbox_normalize(images: Sequence[np.ndarray], input_size, mean=VOC_MEAN, std=VOC_STD, pad_val: int = 128,
                        device="cuda", want_uint8: bool = False, _geometry=None):
    """images: list of uint8 HWC arrays (any sizes).  input_size: int or (h, w).
    -> float32 tensor (B, 3, h, w) = ToTensor(Normalize(Resize(img))) of every image
       [, uint8 tensor (B, h, w, 3) = the padded resized images], geometry [(resize_ratio, du, dl)] per image."""
    th, tw = (input_size, input_size) if isinstance(input_size, int) else (int(input_size[0]), int(input_size[1]))
    device = torch.device(device)
    B = len(images)
    out = torch.empty((B, 3, th, tw), dtype=torch.float32, device=device)
    out_u8 = torch.empty((B, th, tw, 3), dtype=torch.uint8, device=device) if want_uint8 else None
    if B == 0:
        return (out, out_u8, []) if want_uint8 else (out, [])
    if device.type != "cuda":
        raise _lib.PqdetError("letterbox_normalize needs a CUDA device: pqdet_b200 has no CPU path")
    packed, recs, geo = _pack(images, (th, tw), stage=lambda n: _staging(device, n), geometry=_geometry)
    stage = _STAGE[device][0]
    nsrc, nrec = packed.size, recs.nbytes
    rec_at = (nsrc + 7) & ~7                                   # records right behind the pixels, 8-byte aligned
    if rec_at + nrec > stage.numel():
        raise _lib.PqdetError("internal: staging buffer too small")
    stage[rec_at:rec_at + nrec].numpy()[:] = recs.view(np.uint8).reshape(-1)
    dev_buf = stage[:rec_at + nrec].to(device, non_blocking=True)      # ONE host->device copy: pixels + records
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    _STAGE[device][1] = ev
    t_src, t_rec = dev_buf[:nsrc], dev_buf[rec_at:]
    m = (ctypes.c_float * 3)(*[float(np.float32(v)) for v in mean])
    s = (ctypes.c_float * 3)(*[float(np.float32(v)) for v in std])
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.load().pqdet_letterbox_normalize(p(t_src), p(t_rec), B, th, tw, int(pad_val), m, s, p(out), p(out_u8),
                                                     dev_index, ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
               "pqdet_letterbox_normalize")
    return (out, out_u8, geo) if want_uint8 else (out, geo)


def resize_ratio_pad_normalize(images: Sequence[np.ndarray], ratio=1.25, divisor: int = 32, mean=VOC_MEAN, std=VOC_STD,
                               pad_val: int = 128, device="cuda", want_uint8: bool = False):
    """eval_augment_visdrone (dataset/visdrone_sample.py:76-82): ResizeRatio(ratio) -> PadNearestDivisor(pad_val,
    divisor) -> Normalize -> ToTensor.  Every image gets its own canvas (its resized size rounded up to a multiple of
    `divisor`); images that share a canvas size share a launch.
    -> list of float32 tensors (3, H_i, W_i) [, list of uint8 (H_i, W_i, 3)], geometry [(ratio_h, ratio_w, du, dl)]."""
    r = (ratio, ratio) if not isinstance(ratio, (tuple, list)) else tuple(ratio)
    geos = [ratio_pad_geometry(im.shape[:2], r, divisor) for im in images]
    outs, u8s = [None] * len(images), [None] * len(images)
    groups = {}
    for i, g in enumerate(geos):
        groups.setdefault((g[2], g[3]), []).append(i)
    for (ch, cw), idx in groups.items():
        geometry = [(None, geos[i][0], geos[i][1], geos[i][4], geos[i][5]) for i in idx]
        res = letterbox_normalize([images[i] for i in idx], (ch, cw), mean, std, pad_val, device, want_uint8,
                                  _geometry=geometry)
        for j, i in enumerate(idx):
            outs[i] = res[0][j]
            if want_uint8:
                u8s[i] = res[1][j]
    info = [(r[0], r[1], g[4], g[5]) for g in geos]
    return (outs, u8s, info) if want_uint8 else (outs, info)


class Resize:
    """dataset/augment.py:227-259 with the reference's call signature (one image, numpy in / numpy out)."""

    def __init__(self, size, pad_val: int = 128, nopad: bool = False):
        if nopad:
            raise NotImplementedError("nopad=True (train-time multi-scale) is not on the eval path")
        self.size, self.pad_val = size, pad_val

    def __call__(self, img: np.ndarray, bboxes):
        size = self.size() if callable(self.size) else self.size
        _, u8, geo = letterbox_normalize([img], size, pad_val=self.pad_val, want_uint8=True)
        ratio, du, dl = geo[0]
        return u8[0].cpu().numpy(), resize_bboxes(bboxes, ratio, du, dl)
