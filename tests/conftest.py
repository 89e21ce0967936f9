import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the in-tree CUDA library; nvcc cross-compiles without a GPU."""
    import __graft_entry__ as g
    g.build()
    import sndvae_b200 as sv
    return sv
