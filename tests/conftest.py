import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu via gpurun")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden_head():
    return load_golden("head")


@pytest.fixture(scope="session")
def golden_influence():
    return load_golden("influence")


@pytest.fixture(scope="session")
def golden_clusters():
    return load_golden("clusters")


@pytest.fixture(scope="session")
def golden_flow():
    return load_golden("nwnet_flow")


@pytest.fixture(scope="session")
def cuda_lib():
    """The CUDA library on a real GPU.  GPU tests must never pass on a fallback: if the .so is missing
    or the device is not sm_100 this raises instead of skipping."""
    import torch

    from nwhead_b200 import _abi

    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    lib = _abi.load()
    _abi.check(lib.nw_device_check(), "nw_device_check")
    return lib


@pytest.fixture(scope="session")
def golden_env_flow():
    return load_golden("nwnet_env_flow")
