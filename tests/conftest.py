import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def emulated_kernels():
    """Install the oracle-backed CPU emulation of the kernel interface (host-logic tests only)."""
    from swapping_autoencoder_pytorch_b200 import backend
    from tests.cpu_emulation import EmulatedKernels
    prev = backend.set_kernels(EmulatedKernels())
    yield
    backend.set_kernels(prev)


@pytest.fixture
def fp64_default():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(prev)
