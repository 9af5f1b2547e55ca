import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def bundled():
    """Box-filtered + 0.25 m voxelised bundled scan pair (tests/golden/make_fixtures.py)."""
    return dict(np.load(os.path.join(GOLDEN, "bundled_pair.npz")))


@pytest.fixture(scope="session")
def bundled_golden():
    return dict(np.load(os.path.join(GOLDEN, "bundled_pair_golden.npz")))


@pytest.fixture(scope="session")
def spx():
    """The product library through its C-ABI (ctypes).  GPU tests only."""
    import sycl_points_b200 as m
    return m
