import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle.oracle()


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built here with nvcc when stale; never a fallback)."""
    from meshclust_b200 import build
    build.build_lib()
    from meshclust_b200 import api
    return api


@pytest.fixture(scope="session")
def ctx(built_lib):
    c = built_lib.Context(0)
    yield c
    c.close()
