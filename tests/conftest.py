import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def thin(t, cap=32768):
    """Same sub-sampling rule as oracle/make_golden.py:thin."""
    if t.numel() <= cap:
        return t
    flat = t.reshape(-1)
    return flat[:: flat.numel() // cap]


def rel_err(a, b):
    import torch
    a = torch.as_tensor(np.asarray(a)).double() if not hasattr(a, "double") else a.detach().cpu().double()
    b = torch.as_tensor(np.asarray(b)).double() if not hasattr(b, "double") else b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))
