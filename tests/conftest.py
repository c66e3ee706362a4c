"""Shared pytest configuration: the ``gpu`` marker, seeds and the reference's fixture grid
(``tests/conftest.py:25-48`` and ``tests/unit/conftest.py:18-26`` of the reference)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 71892305  # the reference's seed (tests/conftest.py:22)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(SEED)
    yield


@pytest.fixture(autouse=True)
def _tuning_from_env(request):
    """A/B runs of the GPU suite under a kernel-variant knob: MF_TUNING="5=8,4=0" applies
    mf_set_tuning(knob, value) before every GPU test (tests that set knobs restore them to 0)."""
    spec = os.environ.get("MF_TUNING", "")
    if spec and "gpu" in request.keywords:
        from markovflow_b200 import _lib

        for item in spec.split(","):
            k, v = item.split("=")
            _lib.lib().mf_set_tuning(int(k), int(v))
    yield


@pytest.fixture(params=[(3,), (), (2, 1)], ids=["b3", "b0", "b21"])
def batch_shape(request):
    return request.param


@pytest.fixture(params=[1, 3, 5], ids=lambda d: f"D{d}")
def state_dim(request):
    return request.param


@pytest.fixture(params=[1, 3, 5], ids=lambda n: f"N{n}")
def transitions(request):
    return request.param
