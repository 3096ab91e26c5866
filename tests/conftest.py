import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fmb():
    import fmb200
    return fmb200


@pytest.fixture(scope="session")
def gpu(fmb):
    """The CUDA path must be the one that runs: fail loudly (not skip) when it is unavailable under -m gpu."""
    if fmb.device_count() < 1:
        pytest.fail("no CUDA device visible to libfmb200.so: the GPU tests cannot fall back to anything")
    return fmb
