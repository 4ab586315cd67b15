import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))


@pytest.fixture(scope="session")
def golden_r2():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_r2.npz"))


@pytest.fixture(scope="session")
def golden_r3():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_r3.npz"))


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without CUDA."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs CUDA (B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
