"""pytest configuration: the `gpu` marker (tests that need a B200), import path, and
a session fixture that makes sure the in-tree C-ABI library is built."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libcalib_b200.so; builds it (nvcc cross-compiles without a GPU) if absent."""
    from soccernet_calibration_sportlight_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return _lib.LIB_PATH


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
