"""Drop-in check at the outermost boundary: the reference's UNMODIFIED Python package (`clode/*.py`, copied byte for
byte from the reference tree into the git-ignored baseline/_ref by clode_b200.build.build_reference_overlay) runs on
top of this repo's native module, installed under the name the reference imports (`clode.cpp.clode_cpp_wrapper`,
clode/runtime.py:5-15), and the reference's own test-suite (`test/*.py`, also unmodified) passes on it.

CPU: the package imports, every name `clode/__init__.py` re-exports from the native module resolves, the reference's
device-independent tests pass, and the reference's own model files compile verbatim for sm_100a.
GPU: the reference's device tests run as they are (Van der Pol period, ORNL A1 exact solution, aux extents, observers,
OpenCL builtins, trajectories, runtime selection, logger).
"""
import filecmp
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OVERLAY = os.path.join(REPO, "baseline", "_ref")
REFERENCE = os.environ.get("CLODE_REFERENCE_ROOT", "/root/reference")


def _overlay():
    from clode_b200 import build

    if os.path.isdir(os.path.join(REFERENCE, "clode")):
        build.build_reference_overlay()
    if not os.path.exists(os.path.join(OVERLAY, "clode", "__init__.py")):
        pytest.skip("baseline/_ref is not populated (the reference tree is only present in the build container)")
    return OVERLAY


def _run(args, timeout=900):
    env = dict(os.environ, PYTHONPATH=OVERLAY, CLODE_CACHE_DIR=os.path.join(REPO, "clode_b200", "_cubin_cache"))
    return subprocess.run([sys.executable, *args], cwd=OVERLAY, env=env, capture_output=True, text=True, timeout=timeout)


def test_overlay_is_the_unmodified_reference_package():
    root = _overlay()
    if not os.path.isdir(os.path.join(REFERENCE, "clode")):
        pytest.skip("reference tree not present: nothing to compare with")
    names = [n for n in os.listdir(os.path.join(REFERENCE, "clode")) if n.endswith(".py")]
    assert {"solver.py", "features.py", "trajectory.py", "runtime.py", "__init__.py"} <= set(names)
    for n in names:
        assert filecmp.cmp(os.path.join(REFERENCE, "clode", n), os.path.join(root, "clode", n), shallow=False), n
    for n in os.listdir(os.path.join(REFERENCE, "test")):
        if n.endswith((".py", ".cl")):
            assert filecmp.cmp(os.path.join(REFERENCE, "test", n), os.path.join(root, "test", n), shallow=False), n


def test_reference_package_imports_over_this_native_module():
    _overlay()
    r = _run(["-W", "ignore", "-c", "\n".join([
        "import clode, clode.cpp.clode_cpp_wrapper as w, os",
        "assert os.path.realpath(clode.__file__).startswith(os.path.realpath(%r)), clode.__file__" % OVERLAY,
        "assert clode.__version__ == '0.9.0'",
        "names = ['SimulatorBase', 'FeatureSimulatorBase', 'TrajectorySimulatorBase', 'ProblemInfo', 'SolverParams',",
        "         'ObserverParams', 'OpenCLResource', 'CLDeviceType', 'CLVendor', 'DeviceInfo', 'PlatformInfo', 'LogLevel',",
        "         'query_opencl', '_print_opencl', 'get_logger']",
        "missing = [n for n in names if not hasattr(w, n)]",
        "assert not missing, missing",
        "assert clode.Simulator and clode.FeatureSimulator and clode.TrajectorySimulator and clode.Observer.threshold_2",
        "sp = clode.SolverParams(0.1, 0.5, 1e-6, 1e-3, 100, 100, 1); assert sp.max_steps == 100",
        "print('ok', w.__doc__)"])])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "B200" in r.stdout


def test_reference_device_independent_tests_pass_unmodified():
    _overlay()
    r = _run(["-m", "pytest", "-q", "-p", "no:cacheprovider", "-W", "ignore", "test/test_function_converter.py",
              "test/test_clODE_utilities.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("rel, model", [("test/lorenz.cl", "lorenz63"), ("test/van_der_pol_oscillator.cl", "vanderpol"),
                                        ("test/ornl_thompson_a1.cl", "thompson_a1"), ("samples/lactotroph.cl", "lactotroph"),
                                        ("samples/lactotroph_noise.cl", "lactotroph_noise"), ("examples/chay_keizer.cl", "chay_keizer")])
def test_reference_model_files_compile_verbatim(rt, rel, model):
    """the reference's own .cl files, as they are, through NVRTC + the OpenCL-C shim into an sm_100a cubin (no GPU needed)"""
    root = _overlay()
    path = os.path.join(root, rel)
    if not os.path.exists(path):
        pytest.skip(rel + " not in the overlay")
    from clode_b200.models import MODELS

    nv, npar, na, nw = MODELS[model]
    prog = rt.Program(open(path).read(), "seuler" if nw else "dopri5", nv, npar, na, nw, observer="thresh2", min_blocks_per_sm=4)
    cubin, log = rt.compile_program(prog)
    assert cubin[:4] == b"\x7fELF", log


REFERENCE_DEVICE_TESTS = ["test/test_vdp.py", "test/test_features.py", "test/test_ornl_thompson_a1.py", "test/test_aux_values.py",
                          "test/test_opencl_builtins.py", "test/test_observers.py", "test/test_solver.py", "test/test_trajectory.py",
                          "test/test_runtime.py", "test/test_logger.py", "test/test_xpp_parser.py"]


@pytest.mark.gpu
@pytest.mark.parametrize("test_file", REFERENCE_DEVICE_TESTS)
def test_reference_test_suite_passes_unmodified_on_the_gpu(test_file):
    _overlay()
    r = _run(["-m", "pytest", "-q", "-p", "no:cacheprovider", "-W", "ignore", test_file], timeout=1500)
    # 0: everything that ran passed; 5: the file defines no test (test_trajectory.py).  Tests the reference itself marks
    # as skipped (all of test_observers.py and test_solver.py) stay skipped.
    assert r.returncode in (0, 5), r.stdout[-4000:] + r.stderr[-2000:]
    assert " failed" not in r.stdout and " error" not in r.stdout, r.stdout[-2000:]
    if test_file in ("test/test_vdp.py", "test/test_features.py", "test/test_ornl_thompson_a1.py", "test/test_aux_values.py",
                     "test/test_opencl_builtins.py", "test/test_runtime.py", "test/test_logger.py", "test/test_xpp_parser.py"):
        assert " passed" in r.stdout, r.stdout[-2000:]


REFERENCE_EXAMPLES = ["dump_device_performance.py", "spike_counting.py", "Ornstein_Uhlenbeck.py", "observe_sine_curve.py",
                      "find_steady_states.py", "fast_and_slow.py", "phase_response_curve.py", "visualize_events_threshold2.py",
                      "visualize_events_localmax.py", "visualize_events_nhood2.py", "dump_opencl_info.py"]


@pytest.mark.gpu
@pytest.mark.parametrize("script", REFERENCE_EXAMPLES)
def test_reference_example_scripts_run_unmodified_on_the_gpu(script):
    """the reference's own examples/ — Python right-hand sides through its function converter, 2-D parameter grids, every
    observer, stochastic runs, and `dump_device_performance.py`, the only performance script the reference ships — executed as
    they are on the B200 engine (matplotlib is not in the image: a do-nothing stub under tests/stubs stands in for the plots)"""
    root = _overlay()
    path = os.path.join(root, "examples", script)
    if not os.path.exists(path):
        pytest.skip(script + " not in the overlay")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([OVERLAY, os.path.join(HERE, "stubs")]), MPLBACKEND="Agg",
               CLODE_CACHE_DIR=os.path.join(REPO, "clode_b200", "_cubin_cache"))
    r = subprocess.run([sys.executable, "-W", "ignore", path], cwd=OVERLAY, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-2500:]
    if script == "dump_device_performance.py":
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        with open(os.path.join(REPO, "gpurun_out", "reference_dump_device_performance.log"), "w") as f:
            f.write(r.stdout)
        assert "Lorenz system, 1000 RK4 steps" in r.stdout and "131072" in r.stdout
