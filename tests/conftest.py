import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _library_is_built():
    """The built .so is git-ignored: a fresh checkout (or a re-created container) has none.  The tests are about the
    product, which fails loudly without it, so the test session builds it once when nvcc is there (seconds per file;
    cross-compiles without a GPU).  On the GPU box the .so travels with the tree and nothing is rebuilt."""
    import shutil
    from semigcn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        _lib.build_library(verbose=False)
    yield
