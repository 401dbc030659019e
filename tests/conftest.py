import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Oracle (and, where nvcc exists, the CUDA library) are built once per session."""
    import __graft_entry__ as g
    g.build_oracle()
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        g.build_cuda()   # mtime-incremental: a no-op when the library is newer than every source and header
    return g
