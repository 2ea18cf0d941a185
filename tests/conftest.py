import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def S():
    """The product package; the C-ABI library must already be built (no JIT here)."""
    import sdf_viewer_b200 as pkg
    from sdf_viewer_b200 import _lib
    _lib.load()
    return pkg
