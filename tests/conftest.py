import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (REPO, HERE, os.path.join(HERE, "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def rt():
    """the ctypes binding of libclode_rt.so; builds the library on first use"""
    from clode_b200 import build

    build.build_runtime()
    from clode_b200 import _rt

    return _rt


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(HERE, "golden", "golden_ref.npz"))
