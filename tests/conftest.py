import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Builds libcurvis_b200.so and the oracle once per session (no-ops when up to date)."""
    import __graft_entry__ as entry
    entry.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def lib(built):
    from curvis_b200 import _abi
    return _abi.load_library()


@pytest.fixture(scope="session")
def gpu_ctx(lib):
    import curvis_b200 as cv
    return cv.Context([0])
