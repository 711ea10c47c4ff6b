import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {f[:-4]: np.load(os.path.join(d, f), allow_pickle=False) for f in os.listdir(d) if f.endswith(".npz")}


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test was selected but no CUDA device is visible")
    from eonerf_code_b200 import _capi
    _capi.require_device()          # raises unless libeonerf_b200.so is built and the device is sm_100
    return torch.device("cuda:0")
