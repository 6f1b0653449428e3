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
def oracle():
    """The CPU oracle (oracle/libmjp_oracle.so), built on demand."""
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def engine():
    """A handle on cuda:0 through the C ABI (GPU tests only)."""
    from mjpdes_b200 import engine as eng
    h = eng.Handle(device=0, event_hash=True)
    yield h
    h.close()
