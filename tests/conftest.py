import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a box without a device: they are deselected by `-m "not gpu"`
    # and FAIL (not skip) under `-m gpu` if CUDA is missing, so a CPU fallback cannot hide.
    pass


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    return torch.device("cuda:0")
