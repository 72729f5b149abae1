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


def clustered(rs, n, d, n_classes, first_label=1, noise=0.5, background=0.0):
    """Class centroids + noise, L2-normalised (SURVEY.md 8(d) synthetic inputs)."""
    cent = rs.randn(n_classes, d).astype(np.float32)
    lab = rs.randint(0, n_classes, size=n)
    x = cent[lab] + noise * rs.randn(n, d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    labels = (lab + first_label).astype(np.int32)
    if background > 0:
        labels[rs.rand(n) < background] = 0
    return x.astype(np.float32), labels


@pytest.fixture
def rs():
    return np.random.RandomState(12345)
