import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def reference_results():
    return np.load(os.path.join(GOLDEN, "reference_results.npz"))


@pytest.fixture(scope="session")
def oracle_vectors():
    return np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))


@pytest.fixture(scope="session")
def cuda_ready():
    """GPU tests must exercise the native library: fail (not skip) if it cannot load."""
    import torch
    assert torch.cuda.is_available(), "a -m gpu test ran without a CUDA device"
    from irm_motion_planning_b200 import backend
    backend.load_library()
    return True
