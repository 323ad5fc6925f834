import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU parity oracle (oracle/liboracle.so, built on demand).  Checker only."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ip():
    """The product package; the CUDA library must be present (no fallback)."""
    import imagepipe_b200
    imagepipe_b200.lib()
    return imagepipe_b200


@pytest.fixture(scope="session")
def ctx(ip):
    return ip.default_context(0)
