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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: plain `pytest` on a CPU-only machine skips them instead of failing
    (the product has no CPU path).  The torch-fp32 checkers must really compute in fp32: no TF32."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (rangedet_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
