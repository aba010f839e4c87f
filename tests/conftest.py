import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import _bootstrap  # noqa: E402

_bootstrap.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def dfb():
    import del_fem_b200
    return del_fem_b200


@pytest.fixture(scope="session")
def ctx(dfb):
    """one device context per test session (GPU tests only)"""
    c = dfb.Ctx(0)
    yield c
    c.sync()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d
