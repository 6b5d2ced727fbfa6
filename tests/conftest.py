import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference on seeded inputs (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_metrics():
    """Reference decisions / error counts on seeded inputs (tests/golden/make_golden_metrics.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_metrics.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


class Bag:
    """Duck-typed parameter bag (any attribute container works at the boundary)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
