"""Host-side behaviour of the Python front end that needs no GPU: argument validation mirrors the
reference's tests (test/test_runtime.py:6-117), enum / struct surface mirrors clode/__init__.py.
CPU only."""
import os

import pytest

import clode_b200 as clode
from problems import MODELS_DIR

VDP = os.path.join(MODELS_DIR, "vanderpol.cl")


@pytest.mark.parametrize("device_type, vendor, platform_id, device_id, device_ids", [
    ["cpu", "any", 0, None, None], ["cpu", "any", None, 0, None], ["cpu", "any", None, None, [0]],
    ["cpu", None, 0, None, None], ["cpu", None, None, 0, None], ["cpu", None, None, None, [0]],
    [None, "any", 0, None, None], [None, "any", None, 0, None], [None, "any", None, None, [0]],
    [None, None, 0, 0, [0]],
])
def test_incorrect_runtime_config_raises_value_error(device_type, vendor, platform_id, device_id, device_ids):
    dt = clode.CLDeviceType.DEVICE_TYPE_CPU if device_type else None
    vd = clode.CLVendor.VENDOR_ANY if vendor else None
    for cls in (clode.TrajectorySimulator, clode.FeatureSimulator, clode.Simulator):
        with pytest.raises(ValueError):
            cls(src_file=VDP, variables={"x": 1.0, "y": 1.0}, parameters={"mu": 1.0}, num_noise=0,
                stepper=clode.Stepper.dormand_prince, t_span=(0.0, 1000.0), device_type=dt, vendor=vd,
                platform_id=platform_id, device_id=device_id, device_ids=device_ids)


def test_public_names_of_the_reference_package():
    for name in ["Simulator", "FeatureSimulator", "TrajectorySimulator", "Stepper", "Observer", "ObserverOutput",
                 "TrajectoryOutput", "ProblemInfo", "SolverParams", "ObserverParams", "CLDeviceType", "CLVendor",
                 "LogLevel", "set_log_level", "get_log_level", "set_log_pattern", "initialize_runtime", "query_opencl",
                 "print_opencl", "OpenCLResource", "DeviceInfo", "PlatformInfo"]:
        assert getattr(clode, name) is not None
    assert [s.value for s in clode.Stepper] == ["euler", "heun", "rk4", "bs23", "dopri5", "seuler"]
    assert [o.value for o in clode.Observer] == ["basic", "basicall", "localmax", "nhood1", "nhood2", "thresh2"]


def test_every_name_the_reference_package_exports_exists_here():
    """clode/__init__.py (imports + __all__) against this package, including the Python stand-ins of the OpenCL-C
    builtins that Python right-hand sides call (clode.exp, clode.heaviside, clode.pown, ...)"""
    import ast
    import math

    ref = os.path.join(os.environ.get("CLODE_REFERENCE", "/root/reference"), "clode", "__init__.py")
    if os.path.exists(ref):
        names = set()
        for node in ast.walk(ast.parse(open(ref).read())):
            if isinstance(node, ast.ImportFrom):
                names |= {a.asname or a.name for a in node.names}
        missing = sorted(n for n in names if not hasattr(clode, n))
        assert not missing, missing
    assert clode.exp(1.0) == math.e and clode.heaviside(-2.0) == 0.0 and clode.heaviside(3.0) == 1.0
    assert clode.pown(2.0, 3) == 8.0 and clode.rootn(27.0, 3) == pytest.approx(3.0) and clode.sinpi(0.5) == pytest.approx(1.0)
    from clode_b200 import cospi, exp10, rsqrt  # noqa: F401  (importable by name)
    assert set(clode._BUILTINS) <= set(clode.__all__)


def test_struct_defaults_match_the_reference_binding():
    sp = clode.SolverParams()  # clode/cpp/CLODEpython.cpp:229-235
    assert (sp.dt, sp.dtmax, sp.abstol, sp.reltol, sp.max_steps, sp.max_store, sp.nout) == (0.1, 0.5, 1e-6, 1e-3, 1000000, 1000000, 1)
    op = clode.ObserverParams()  # clode/cpp/CLODEpython.cpp:302-313
    assert (op.e_var_ix, op.f_var_ix, op.max_event_count, op.max_event_timestamps) == (0, 0, 100, 0)
    assert (op.nhood_radius, op.x_up_threshold, op.x_down_threshold) == (0.05, 0.2, 0.2)
    pi = clode.ProblemInfo(VDP, ["x", "y"], ["mu"])
    assert (pi.num_var, pi.num_par, pi.num_aux, pi.num_noise) == (2, 1, 0, 1)
    pi.aux = ["a", "b"]
    assert pi.num_aux == 2 and "problem_info" in repr(pi)


def test_log_level_round_trip():
    previous = clode.get_log_level()
    clode.set_log_level(clode.LogLevel.trace)
    assert clode.get_log_level() == clode.LogLevel.trace
    clode.set_log_level(clode.LogLevel.off)
    assert clode.get_log_level() == clode.LogLevel.off
    clode.set_log_level(previous)


def test_source_arguments_are_validated_like_the_reference():
    """clode/solver.py:228-250"""
    with pytest.raises(ValueError, match="Cannot specify both"):
        clode.Simulator(variables={"x": 0.0}, parameters={"a": 1.0}, src_file="a.cl", rhs_equation=lambda *a: None)
    with pytest.raises(ValueError, match="Must specify either"):
        clode.Simulator(variables={"x": 0.0}, parameters={"a": 1.0})

    def untyped(t, x, p, dx, aux, w):  # the converter refuses it before any device work starts
        dx[0] = -x[0]

    with pytest.raises(TypeError, match="must have a return type"):
        clode.Simulator(variables={"x": 0.0}, parameters={"a": 1.0}, rhs_equation=untyped)


def test_missing_gpu_is_a_loud_error_not_a_fallback(rt):
    try:
        have = rt.device_count() > 0
    except rt.RtError:
        have = False
    if have:
        pytest.skip("a CUDA driver is present")
    with pytest.raises(RuntimeError):
        clode.Simulator(src_file=VDP, variables={"x": 1.0, "y": 1.0}, parameters={"mu": 1.0})
