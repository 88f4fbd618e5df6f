import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_close(a, b, rtol=1e-5, scale=None):
    """|a-b| <= rtol * max(|b|, scale) elementwise.  `scale` is the natural magnitude of the
    quantity (e.g. the network input size for box coordinates, whose fp32 ulp sets the floor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if scale is None:
        scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-30)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), scale)))
