import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "sdfa-2019_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def flame():
    from deformation import workloads as W
    V, F, nfv, nft = W.load_flame()
    return dict(V=V, F=F, nfv=nfv, nft=nft, tol=1e-6 * W.bbox_diag(V))


@pytest.fixture(scope="session")
def golden_flame():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_flame.npz"))


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))
