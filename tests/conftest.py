import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (it is git-ignored): build it once (nvcc cross-compiles
    without a GPU, ~30 s) instead of failing every test that loads it."""
    if not os.path.exists(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "libebncuda.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def ebn():
    """The product package, initialised on cuda:0. Fails loudly without the CUDA library/GPU."""
    import ebn_b200
    ebn_b200.init()
    yield ebn_b200
    ebn_b200.shutdown()
