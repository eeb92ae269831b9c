import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """A fresh checkout has no libqcb200.so yet (built artefacts are not in the history): compile it once (nvcc cross-compiles
    without a GPU).  An existing library is used as it is - `__graft_entry__.build()` is what rebuilds a stale one."""
    from qclojure_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()


GOLDEN = os.path.join(ROOT, "tests", "golden")
