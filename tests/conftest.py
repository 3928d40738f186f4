import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def small_ice():
    return load_golden("small_ice.npz")


@pytest.fixture(scope="session")
def small_kr():
    return load_golden("small_kr.npz")


@pytest.fixture(scope="session")
def gm12878():
    return load_golden("gm12878.npz")


@pytest.fixture(scope="session")
def kr_partial():
    return load_golden("kr_partial.npz")
