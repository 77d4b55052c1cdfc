import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never be silently skipped on the GPU box: they fail if CUDA is missing
    when selected with -m gpu; on CPU boxes they are simply deselected by -m 'not gpu'."""


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    import vadx
    vadx.lib.load()
    vadx.lib.require_device()
    return torch.device("cuda:0")
