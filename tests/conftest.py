import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emul():
    """ctypes handle on the CPU replay of the kernels' host/device headers (tests/emul)."""
    import ctypes

    from tests.emul.build import build

    lib = ctypes.CDLL(build())
    lib.emul_scan.restype = ctypes.c_int64
    return lib
