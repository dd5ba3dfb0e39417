import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (dev container only)")


@pytest.fixture(scope="session")
def ctx():
    """One C-ABI context for the whole GPU session (one context per process and GPU)."""
    from nbed_b200.backend import B200Context

    c = B200Context(0)
    yield c
    c.close()
