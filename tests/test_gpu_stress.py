"""-m gpu: bounded slices of the randomised sweeps (tests/stress_gpu*.py hold the case generators; run by hand they
go on for hundreds of cases).  Here ~30 seeded cases of each run on every `pytest -m gpu`, split into chunks so a
failure names its neighbourhood."""
import numpy as np
import pytest
import torch

import stress_gpu
import stress_gpu_decode
import stress_gpu_train

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunk", range(3))
def test_eval_path_sweep(chunk):
    from pqdet_b200 import config
    old = config.nms_semantics
    rng = np.random.default_rng(1000 + chunk)
    try:
        kept = sum(stress_gpu.one_case(rng, 10 * chunk + i) for i in range(10))
    finally:
        config.nms_semantics = old
    assert kept > 0


@pytest.mark.parametrize("chunk", range(3))
def test_training_path_sweep(chunk):
    rng = np.random.default_rng(2000 + chunk)
    for i in range(10):
        stress_gpu_train.train_case(rng, 10 * chunk + i)


def test_section_8f_components_sweep():
    rng = np.random.default_rng(3000)
    for i in range(12):
        stress_gpu_train.misc_case(rng, i)


@pytest.mark.parametrize("chunk", range(2))
def test_persistent_vs_general_decode_and_head_conv_sweep(chunk):
    rng = np.random.default_rng(4000 + chunk)
    torch.manual_seed(4000 + chunk)
    for i in range(16):
        (stress_gpu_decode.decode_case if i % 2 == 0 else stress_gpu_decode.headconv_case)(rng)
