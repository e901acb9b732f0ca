import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "collaborative-gan-sampling_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


@pytest.fixture(scope="session")
def cgs_lib():
    """libcgs.so built in-tree (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, PKG)
    import build as cgs_build  # collaborative-gan-sampling_b200/build.py
    cgs_build.build_lib()
    from cgs import lib
    return lib.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
