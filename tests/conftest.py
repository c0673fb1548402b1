import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lk():
    """The product library through its C ABI."""
    import loki_b200
    return loki_b200.load()


@pytest.fixture(scope="session")
def ok():
    """The CPU oracle (test infrastructure)."""
    import oracle_binding
    return oracle_binding.load()


@pytest.fixture()
def strict(lk):
    old = lk.lk_set_strict(1)
    yield
    lk.lk_set_strict(old)


@pytest.fixture()
def fast(lk):
    old = lk.lk_set_strict(0)
    yield
    lk.lk_set_strict(old)
