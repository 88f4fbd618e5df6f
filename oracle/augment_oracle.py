"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (pqdet_b200/).

CPU restatement of the eval pre-processing chain of eleflea/PQDet: augment.Resize (dataset/augment.py:227-259),
augment.Normalize (:206-215), augment.ToTensor (:390-398).  The arithmetic that matters lives in a third-party
dependency of the reference: cv2.resize(INTER_LINEAR) on 8-bit images (opencv-python, requirements.txt; 4.13.0
installed) = OpenCV's fixed-point bilinear resize (modules/imgproc/src/resize.cpp: 11-bit coefficients,
INTER_RESIZE_COEF_BITS; x taps reset at the borders, y taps not - their two source rows are clipped instead).
Pinned bit-exactly against cv2 itself (tests/test_oracle_vs_reference.py) and against the reference's own classes
(tests/golden/letterbox.npz)."""
from __future__ import annotations

import numpy as np


def _taps(dn: int, sn: int, scale: float, reset_at_border: bool):
    ofs = np.zeros(dn, np.int64)
    c = np.zeros((dn, 2), np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if reset_at_border:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= sn - 1:
                f, s = np.float32(0), sn - 1
        ofs[d] = s
        c[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        c[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return ofs, c


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images."""
    sh, sw = img.shape[:2]
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)
    xo, xa = _taps(dw, sw, scale_x, True)
    yo, ya = _taps(dh, sh, scale_y, False)
    src = img.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    H = src[:, xo, :] * xa[None, :, 0, None] + src[:, x1, :] * xa[None, :, 1, None]
    y0, y1 = np.clip(yo, 0, sh - 1), np.clip(yo + 1, 0, sh - 1)
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (H[y0] >> 4)) >> 16) + ((b1 * (H[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_letterbox(img: np.ndarray, target_hw, pad_val: int = 128):
    """augment.Resize.__call__ without the bboxes: -> padded uint8 image, (resize_ratio, du, dl)."""
    target_h, target_w = target_hw
    img_h, img_w = img.shape[:2]
    resize_ratio = min(target_w / img_w, target_h / img_h)
    resize_w = round(resize_ratio * img_w)
    resize_h = round(resize_ratio * img_h)
    resized = resize_linear_u8(img, resize_w, resize_h)
    dl = (target_w - resize_w) // 2
    dr = target_w - resize_w - dl
    du = (target_h - resize_h) // 2
    dd = target_h - resize_h - du
    padded = np.pad(resized, ((du, dd), (dl, dr), (0, 0)), 'constant', constant_values=pad_val)
    return padded, (resize_ratio, du, dl)


def normalize_to_chw(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """augment.Normalize + augment.ToTensor: float32 (3, H, W)."""
    mean, std = np.array(mean, dtype=np.float32), np.array(std, dtype=np.float32)
    img = img_u8.astype(np.float32, copy=False)
    img = (img / 255. - mean) / std
    return np.transpose(img, (2, 0, 1)).astype(np.float32)


def resize_ratio_pad(img: np.ndarray, ratio=(1.25, 1.25), pad_val: int = 128, divisor: int = 32):
    """augment.ResizeRatio (dataset/augment.py:261-273) + augment.PadNearestDivisor (:275-298) without the bboxes."""
    from math import ceil
    target_h, target_w = [round(a * b) for a, b in zip(ratio, img.shape[:2])]
    resized = resize_linear_u8(img, target_w, target_h)
    img_h, img_w = resized.shape[:2]
    th, tw = int(ceil(img_h / divisor) * divisor), int(ceil(img_w / divisor) * divisor)
    dl = (tw - img_w) // 2
    dr = tw - img_w - dl
    du = (th - img_h) // 2
    dd = th - img_h - du
    return np.pad(resized, ((du, dd), (dl, dr), (0, 0)), 'constant', constant_values=pad_val), (du, dl)
