import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def palib():
    """The built C-ABI library (host-only entry points work without a GPU)."""
    from peleanalysis_b200 import build, capi
    build.build()
    return capi.lib()


@pytest.fixture(scope="session")
def gpu(palib):
    from peleanalysis_b200 import capi
    capi.init(0)       # raises PaError if there is no device: GPU tests must fail loudly, never fall back
    return capi
