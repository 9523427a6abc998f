import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test session (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "adaptive-graph-convolutional-network_b200"))
    import build_ext
    build_ext.build()
    yield
