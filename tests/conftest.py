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


def _cuda_devices():
    try:
        from clode_b200 import _rt, build

        build.build_runtime()
        return _rt.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device skips the gpu-marked tests instead of failing them
    (CLODE_PRECOMPILE=1 keeps them: that mode compiles their programs and skips at Sim creation)."""
    if os.environ.get("CLODE_PRECOMPILE") == "1" or not any("gpu" in it.keywords for it in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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


@pytest.fixture(autouse=True, scope="session")
def _precompile_only():
    """CLODE_PRECOMPILE=1 python -m pytest tests -m gpu  (on a machine WITHOUT a GPU): every program a GPU test
    would build through clode_b200._rt.Sim is compiled into the cubin cache instead (the first occupancy
    candidates the runtime would try) and the test is skipped — the GPU box then finds the cubins and skips the JIT."""
    if os.environ.get("CLODE_PRECOMPILE") != "1":
        yield
        return
    import dataclasses

    from clode_b200 import _rt

    def init(self, prog, device=0):
        self._h = None
        blocks = [prog.min_blocks_per_sm] if prog.min_blocks_per_sm else ([8, 5, 4] if prog.single_precision else [5, 4])
        for m in blocks:
            _rt.compile_program(dataclasses.replace(prog, min_blocks_per_sm=m))
        # the runtime moves the extents of a multi-variable observer to shared memory when the features kernel spills
        # (clode_sim_build); which way it goes is only known on the GPU, so both variants are put into the cache
        if prog.observer != "basic" and (prog.kernels & _rt.KERNEL_FEATURES) and not prog.observer_in_shared and "CLODE_EXT_SMEM" not in os.environ:
            os.environ["CLODE_EXT_SMEM"] = "1"
            try:
                for m in blocks:
                    _rt.compile_program(dataclasses.replace(prog, min_blocks_per_sm=m))
            finally:
                del os.environ["CLODE_EXT_SMEM"]
        pytest.skip("precompiled")

    original = _rt.Sim.__init__
    _rt.Sim.__init__ = init
    yield
    _rt.Sim.__init__ = original
