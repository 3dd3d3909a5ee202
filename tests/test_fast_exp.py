"""exp() of production double builds (clode_b200/csrc/device/fast_exp.cuh).

CPU: the device function compiled as host C++ against the 80-bit expl of the host (<= 0.53 ulp, special values).
GPU: the same function inside a program, against numpy's longdouble exp (test_cuda_fast_exp_*)."""
import os
import subprocess

import numpy as np
import pytest

from clode_b200 import _rt

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# x' = a, aux0 = exp(x), aux1 = native library exp for reference
RHS = """void getRHS(const realtype t, const realtype x_[], const realtype p_[], realtype dx_[], realtype aux_[], const realtype w_[]) {
    dx_[0] = p_[0];
    aux_[0] = exp(x_[0]);
}
"""


def test_fast_exp_error_bound_on_the_host(tmp_path):
    exe = str(tmp_path / "fast_exp_check")
    subprocess.run(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", f"-I{REPO}/clode_b200/csrc/device",
                    f"{REPO}/tests/emu/fast_exp_check.cpp", "-o", exe], check=True)
    out = subprocess.run([exe, "3000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "bad=0" in out.stdout


def test_fast_exp_is_compiled_into_production_double_only():
    base = dict(rhs_source=RHS, stepper="euler", n_var=1, n_par=1, n_aux=1, kernels=_rt.KERNEL_TRAJECTORY)
    src = _rt.program_source(_rt.Program(**base))
    assert "clode_fast_exp" in src and "-DCLODE_LIBRARY_EXP" not in src
    assert "-DCLODE_LIBRARY_EXP" in _rt.program_source(_rt.Program(**base, library_exp=True))
    for kw in (dict(), dict(library_exp=True), dict(bit_exact=True), dict(single_precision=True)):
        cubin, _ = _rt.compile_program(_rt.Program(**base, **kw))
        assert cubin[:4] == b"\x7fELF"


def _exp_on_gpu(library_exp):
    n, rows = 4096, 200
    prog = _rt.Program(RHS, "euler", 1, 1, 1, kernels=_rt.KERNEL_TRAJECTORY, library_exp=library_exp)
    sim = _rt.Sim(prog)
    sim.set_solver_params(dt=1.0, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=rows, max_store=rows + 1, nout=1)
    sim.set_tspan(0.0, float(rows))
    rng = np.random.default_rng(5)
    x0 = rng.uniform(-700.0, -650.0, n)
    slope = rng.uniform(0.0, 7.0, n)          # x sweeps [-700, 700] over the 200 stored points
    sim.set_problem(x0, slope)
    sim.seed_rng(1)
    sim.trajectory()
    tr = sim.get_trajectory()
    keep = int(tr["n_stored"].min()) + 1
    x = tr["x"].reshape(-1, n)[:keep].copy()
    aux = tr["aux"].reshape(-1, n)[:keep].copy()
    sim.close()
    return x.ravel(), aux.ravel()


@pytest.mark.gpu
def test_cuda_fast_exp_within_one_ulp_of_longdouble():
    x, y = _exp_on_gpu(False)
    ref = np.exp(x.astype(np.longdouble))
    ok = np.isfinite(y) & (ref > np.finfo(np.float64).tiny)
    ulp = np.spacing(ref[ok].astype(np.float64)).astype(np.longdouble)
    err = np.abs(y[ok].astype(np.longdouble) - ref[ok]) / ulp
    assert ok.sum() > 500000
    assert float(err.max()) <= 0.6, float(err.max())      # 0.51 measured on the host; tolerance: 0.6 ulp


@pytest.mark.gpu
def test_cuda_fast_exp_against_the_library_exp():
    x, fast = _exp_on_gpu(False)
    x2, lib = _exp_on_gpu(True)
    assert np.array_equal(x, x2)
    ok = np.isfinite(lib) & (lib > np.finfo(np.float64).tiny)
    assert np.all(np.abs(fast[ok] - lib[ok]) <= 2 * np.spacing(lib[ok]))   # 0.52 + 1 ulp, in units of the larger spacing
    assert np.array_equal(fast[~ok], lib[~ok])                             # overflow / underflow: the library's own values


DIV_RHS = """void getRHS(const realtype t, const realtype x_[], const realtype p_[], realtype dx_[], realtype aux_[], const realtype w_[]) {
    dx_[0] = p_[0];
    aux_[0] = div_norm(x_[0], p_[1]);   /* the engine's error-norm division (steppers.cuh), visible to the RHS text */
    aux_[1] = div_nr(x_[0], p_[1]);     /* the bookkeeping division: two Newton steps */
}
"""


@pytest.mark.gpu
def test_cuda_engine_divisions_accuracy():
    """production double: div_norm (one Newton step on the SFU reciprocal, used only inside the step-size controller's
    error norm) is within 1e-11 of the IEEE quotient, div_nr (two steps, running means) within 2 ulp"""
    n, rows = 4096, 60
    prog = _rt.Program(DIV_RHS, "euler", 1, 2, 2, kernels=_rt.KERNEL_TRAJECTORY)
    sim = _rt.Sim(prog)
    sim.set_solver_params(dt=1.0, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=rows, max_store=rows + 1, nout=1)
    sim.set_tspan(0.0, float(rows))
    rng = np.random.default_rng(11)
    x0 = rng.uniform(-5.0, 5.0, n)
    pars = np.concatenate([rng.uniform(0.01, 3.0, n), 10.0 ** rng.uniform(-12, 12, n)])
    sim.set_problem(x0, pars)
    sim.seed_rng(1)
    sim.trajectory()
    tr = sim.get_trajectory()
    keep = int(tr["n_stored"].min()) + 1
    x = tr["x"].reshape(-1, n)[:keep]
    aux = tr["aux"].reshape(-1, 2, n)[:keep]
    sim.close()
    want = x / pars[n:][None, :]
    ok = want != 0.0
    rel = np.abs(aux[:, 0][ok] - want[ok]) / np.abs(want[ok])
    assert rel.max() < 1e-11, rel.max()
    assert np.all(np.abs(aux[:, 1][ok] - want[ok]) <= 2 * np.spacing(np.abs(want[ok])))
