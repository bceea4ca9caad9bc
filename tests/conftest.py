import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu through gpurun)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test started without a CUDA device (run `-m 'not gpu'` on CPU boxes)")
    from vlmfw_b200 import _lib

    _lib.load()  # loud failure if the .so was not built / shipped
    assert _lib.device_ok(), _lib.last_error()
    return torch.device("cuda:0")
