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
def oracle():
    """The CPU oracle (test infrastructure only; see oracle/psoracle_impl.h)."""
    from oracle import psoracle
    psoracle.build()
    return psoracle


@pytest.fixture(scope="session")
def ps():
    import powerspectra_jl_b200
    return powerspectra_jl_b200


def parity_error(test, ref, floor=1e-30):
    """North-star criterion: max over rows of |test-ref|/|ref| on every entry with
    |ref| > floor * max|row| (BASELINE.json)."""
    test = np.asarray(test)
    ref = np.asarray(ref)
    rowmax = np.max(np.abs(ref), axis=1, keepdims=True)
    sel = np.abs(ref) > floor * rowmax
    if not sel.any():
        return 0.0
    rel = np.zeros_like(ref)
    rel[sel] = np.abs(test[sel] - ref[sel]) / np.abs(ref[sel])
    return float(rel.max())
