import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gold_tiny():
    return np.load(os.path.join(GOLD, "tiny_dynamic.npz"))


@pytest.fixture(scope="session")
def gold_full():
    return np.load(os.path.join(GOLD, "full3b.npz"))


@pytest.fixture(scope="session")
def gold_thr():
    return np.load(os.path.join(GOLD, "thresholds.npz"))
