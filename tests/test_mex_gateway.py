"""The MATLAB entry points (clode_b200/csrc/matlab: clODEmex, clODEfeaturesmex, clODEtrajectorymex, queryOpenCL)
built against the in-tree mex stub and driven by tests/emu/mex_harness.cpp with the reference's command protocol
(matlab/clODEmex.cpp:59-90).  CPU: they compile, link, validate arguments and fail loudly without a GPU.  GPU: the
same calls run a Lorenz ensemble and the results equal the C ABI path."""
import os
import subprocess

import numpy as np
import pytest

from clode_b200 import build
from clode_b200.models import rhs_path

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MATLAB = os.path.join(REPO, "clode_b200", "csrc", "matlab")
HOST = os.path.join(REPO, "clode_b200", "csrc", "host")
MEX = ["clODEmex", "clODEfeaturesmex", "clODEtrajectorymex"]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    build.build_all()
    out = tmp_path_factory.mktemp("mex")
    host_srcs = [os.path.join(HOST, f) for f in ("CLODE.cpp", "CLODEfeatures.cpp", "CLODEtrajectory.cpp", "OpenCLResource.cpp")]
    exes = []
    for which, name in enumerate(MEX):
        exe = str(out / name)
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", f"-DWHICH={which}", f"-I{MATLAB}/stub", f"-I{MATLAB}", f"-I{HOST}",
                        f"-I{REPO}/include", os.path.join(MATLAB, name + ".cpp"), os.path.join(MATLAB, "stub", "mex_stub.cpp"),
                        os.path.join(REPO, "tests", "emu", "mex_harness.cpp"), *host_srcs, "-o", exe,
                        f"-L{REPO}/clode_b200", "-lclode_rt", f"-Wl,-rpath,{REPO}/clode_b200"], check=True)
        exes.append(exe)
    return exes


def _run(exe, mode):
    r = subprocess.run([exe, rhs_path("lorenz63"), mode], capture_output=True, text=True)
    lines = dict(l.split(" ", 1) for l in r.stdout.splitlines() if " " in l)
    return r, lines


def test_query_opencl_entry_point_compiles():
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", f"-I{MATLAB}/stub", f"-I{HOST}", f"-I{REPO}/include",
                    os.path.join(MATLAB, "queryOpenCL.cpp")], check=True)


@pytest.mark.parametrize("which", [0, 1, 2])
def test_mex_entry_points_validate_arguments_and_fail_loudly_without_a_gpu(harness, which):
    import clode_b200._rt as rt
    try:
        has_gpu = rt.device_count() > 0
    except rt.RtError:
        has_gpu = False
    r, lines = _run(harness[which], "cpu")
    assert lines["nohandle"] == "clODE:args" and lines["badhandle"] == "clODE:handle" and lines["nocommand"] == "clODE:args"
    if not has_gpu:
        assert r.returncode == 0
        assert lines["new-error"].startswith("clODE:runtime"), r.stdout     # no CPU fallback: the constructor raises
        assert "handle" not in lines


@pytest.mark.gpu
@pytest.mark.parametrize("which", [0, 1, 2])
def test_mex_entry_points_run_an_ensemble(harness, which, rt):
    r, lines = _run(harness[which], "gpu")
    assert r.returncode == 0, r.stdout + r.stderr
    assert lines["handle"].startswith("1 locks 1") and lines["deleted"].startswith("1 locks 0")
    assert lines["steppers"] == "6" and lines["unknown"] == "clODE:command" and "setnpts is ignored" in lines["warning"]
    assert lines["tspan"].startswith("2 x 1 : 0 5")
    # the same ensemble through the C ABI
    from clode_b200.models import MODELS, rhs_source
    n = 64
    pars = np.concatenate([5.0 + 20.0 * np.arange(n) / (n - 1), np.full(n, 10.0), np.full(n, 8.0 / 3.0)])
    prog = rt.Program(rhs_source("lorenz63"), "dopri5", *MODELS["lorenz63"], kernels=rt.KERNEL_TRANSIENT)
    sim = rt.Sim(prog)
    sim.set_solver_params(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, max_steps=100000, max_store=50, nout=1)
    sim.set_tspan(0.0, 5.0)
    sim.set_problem(np.ones(3 * n), pars)
    sim.seed_rng(1)
    sim.transient()
    want = sim.get_xf()
    sim.close()
    shape, values = lines["xf"].split(" :")
    assert shape == "1 x 192"                                       # a row vector, as the reference returns it
    assert np.array_equal(np.array(values.split(), dtype=float), want[:8])
    if which == 1:
        assert lines["nfeatures"] == "6" and lines["F"].startswith("384 x 1") and lines["featurenames"].startswith("6 observers 6")
    if which == 2:
        assert lines["nstored"].startswith("64 x 1") and lines["x"].startswith(f"{50 * 3 * n} x 1")
