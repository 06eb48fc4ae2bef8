import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def drfe():
    import drfe as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as m
    m.lib()
    return m


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def sort_rows(a):
    a = np.asarray(a)
    return a[np.lexsort((a[:, 2], a[:, 0], a[:, 1]))] if len(a) else a


KP_FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")
