import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ref_ops(golden_dir):
    """Outputs of the reference's own compiled CUDA kernels (oracle/gen_golden_gpu.py)."""
    import numpy as np

    p = os.path.join(golden_dir, "ref_ops.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/ref_ops.npz not generated yet")
    return np.load(p)
