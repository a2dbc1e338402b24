import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the `gpu` tier instead of failing in it."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_A():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "nvf_A.npz"))


@pytest.fixture(scope="session")
def golden_B():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "nvf_B.npz"))


@pytest.fixture(scope="session")
def gpu():
    """The product binding on a CUDA device (the `-m gpu` tier)."""
    import torch
    from nvfpcc_b200 import _lib
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return _lib.cuda_binding()
