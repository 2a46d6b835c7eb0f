"""pytest configuration: `gpu` marker, import paths, one in-tree build per session."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "falcon-ttdforgnns_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ttg_lib():
    """The C-ABI library; built in-tree if nvcc is here, otherwise it must already exist."""
    import shutil
    import _ttg
    if not os.path.exists(_ttg.LIB_PATH) and shutil.which("nvcc"):
        sys.path.insert(0, PKG)
        import build as ttg_build
        ttg_build.build()
    return _ttg.lib()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "tt_cases.npz"))
    cases = {}
    for k in z.files:
        name, field = k.split("/", 1)
        cases.setdefault(name, {})[field] = z[k]
    return cases
