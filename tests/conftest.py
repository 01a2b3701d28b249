import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pre3():
    """The product package (its name starts with a digit, hence import_module)."""
    return importlib.import_module("3pre_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("3pre_b200.synth")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "siftmatch_ref.npz"))


@pytest.fixture(scope="session")
def ctx(pre3):
    """One libpre3 context on cuda:0 for the whole GPU session."""
    c = pre3.Context(0)
    yield c
    c.close()
