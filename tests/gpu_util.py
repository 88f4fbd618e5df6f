"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch

from oracle import pqdet_oracle as po
from pqdet_b200 import synth


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def assert_same_detections(got: np.ndarray, want: np.ndarray, ties_unordered: bool = False, what=""):
    """Bit-exact comparison of (K,6) detection lists.  With ties_unordered (torchvision's vanilla path
    re-sorts with an unstable sort) rows of equal score may come in any order."""
    got = np.asarray(got).reshape(-1, 6) if np.asarray(got).size else np.zeros((0, 6), np.float32)
    want = np.asarray(want).reshape(-1, 6) if np.asarray(want).size else np.zeros((0, 6), np.float32)
    assert got.shape == want.shape, "%s: kept %d vs %d" % (what, got.shape[0], want.shape[0])
    if not ties_unordered:
        assert np.array_equal(got, want), "%s: rows differ" % what
        return
    assert np.array_equal(got[:, 4], want[:, 4]), "%s: scores differ" % what
    assert sorted(map(bytes, got)) == sorted(map(bytes, want)), "%s: row sets differ" % what


def eval_chain_oracle(heads_np, strides, C, input_size, orig, kind, thr, iou, device_sem, mode="auto",
                      decoded=None):
    """Reference sequence on OUR decoded tensor (bit-exactness is defined on identical boxes):
    recover -> per-image torch_nms, all by the oracle."""
    pred = decoded if decoded is not None else po.detect(heads_np, C, strides)
    rec = po.recover(pred, input_size, orig, kind)
    return [po.torch_nms(rec[b], thr, iou, device=device_sem, mode=mode, return_index=True)
            for b in range(rec.shape[0])]
