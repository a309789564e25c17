import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    from ninpol_b200 import build as nb
    from ninpol_b200 import _capi
    nb.build()      # a no-op when the library is newer than every source; never test a stale binary
    return _capi.load_library()


@pytest.fixture(scope="session", autouse=True)
def _library_matches_sources():
    """Every test session (CPU or GPU box) runs against a library built from the sources in the tree: the
    build is content-hashed, so this is a no-op unless a .cu / .cuh file changed since the last build."""
    from ninpol_b200 import build as nb
    nb.build()
