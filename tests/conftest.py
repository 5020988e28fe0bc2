import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_PRESENT = os.path.exists("/root/reference/lmcma_path_planner/src/lmcma.cpp")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "lmcma_reference.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_maps():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "maps_2d.npz")))


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


def weighted_sphere(X):
    X = np.atleast_2d(X)
    return np.sum((1.0 + np.arange(X.shape[1])) * X * X, axis=1)
