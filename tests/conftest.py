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


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle.oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own C++ (oracle/_ref/librd_ref.so); None when neither /root/reference nor a
    prebuilt copy exists."""
    import oracle
    return oracle.reference()


def rel_err(a, b):
    """Normwise relative error max|a-b| / max|b| (the 1e-3 'rel' of BASELINE.json's north_star)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
