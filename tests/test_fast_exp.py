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


def test_branch_free_exp_error_bound_on_the_host(tmp_path):
    """the 2048-entry, branch-free variant (CLODE_EXP_2K; CLODE_BRANCHLESS=1 builds): <= 1.1 ulp, IEEE answers for
    overflow / underflow / Inf / NaN / huge arguments, every table entry built at kernel entry correctly rounded"""
    exe = str(tmp_path / "fast_exp2k_check")
    subprocess.run(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-DCLODE_EXP_2K", f"-I{REPO}/clode_b200/csrc/device",
                    f"{REPO}/tests/emu/fast_exp_check.cpp", "-o", exe], check=True)
    out = subprocess.run([exe, "3000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "bad=0" in out.stdout


def test_polar_scale_error_bound_on_the_host(tmp_path):
    """device/fast_polar.cuh, sqrt(-2 log q / q) of the polar method in production double builds of the stochastic stepper:
    <= 3 ulp against 80-bit arithmetic over (0, 1), including q -> 1 (where a naive log cancels) and q -> 0"""
    exe = str(tmp_path / "fast_polar_check")
    subprocess.run(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", f"-I{REPO}/clode_b200/csrc/device",
                    f"{REPO}/tests/emu/fast_polar_check.cpp", "-o", exe], check=True)
    out = subprocess.run([exe, "2000000"], capture_output=True, text=True)
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
def test_cuda_fast_exp_within_one_ulp_of_longdouble(monkeypatch):
    monkeypatch.setenv("CLODE_BRANCHLESS", "0")   # the 128-entry hi/lo variant (CLODE_BRANCHLESS=1, the default, selects the 2048-entry one)
    x, y = _exp_on_gpu(False)
    ref = np.exp(x.astype(np.longdouble))
    ok = np.isfinite(y) & (ref > np.finfo(np.float64).tiny)
    ulp = np.spacing(ref[ok].astype(np.float64)).astype(np.longdouble)
    err = np.abs(y[ok].astype(np.longdouble) - ref[ok]) / ulp
    assert ok.sum() > 500000
    assert float(err.max()) <= 0.6, float(err.max())      # 0.51 measured on the host; tolerance: 0.6 ulp


@pytest.mark.gpu
def test_cuda_branch_free_exp_within_1p2_ulp_of_longdouble(monkeypatch):
    monkeypatch.setenv("CLODE_BRANCHLESS", "1")
    x, y = _exp_on_gpu(False)
    ref = np.exp(x.astype(np.longdouble))
    ok = np.isfinite(y) & (ref > np.finfo(np.float64).tiny)
    ulp = np.spacing(ref[ok].astype(np.float64)).astype(np.longdouble)
    err = np.abs(y[ok].astype(np.longdouble) - ref[ok]) / ulp
    assert ok.sum() > 500000
    assert float(err.max()) <= 1.2, float(err.max())      # 1.06 measured on the host


@pytest.mark.gpu
@pytest.mark.parametrize("branchless", ["0", "1"])
def test_cuda_fast_exp_against_the_library_exp(monkeypatch, branchless):
    monkeypatch.setenv("CLODE_BRANCHLESS", branchless)
    x, fast = _exp_on_gpu(False)
    x2, lib = _exp_on_gpu(True)
    assert np.array_equal(x, x2)
    ok = np.isfinite(lib) & (lib > np.finfo(np.float64).tiny)
    assert np.all(np.abs(fast[ok] - lib[ok]) <= 3 * np.spacing(lib[ok]))   # 0.52 (1.06) + 1 ulp, in units of the larger spacing
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


# one instance per operand pair: aux0 = 1 / x, aux1 = a / x, aux2 = exp(a) at the INITIAL point (row 0 of the trajectory)
OPS_RHS = """void getRHS(const realtype t, const realtype x_[], const realtype p_[], realtype dx_[], realtype aux_[], const realtype w_[]) {
    dx_[0] = RCONST(0.0);
    aux_[0] = RCONST(1.0) / x_[0];
    aux_[1] = p_[0] / x_[0];
    aux_[2] = exp(p_[0]);
}
"""


def _ops_on_gpu(x, a):
    n = x.size
    prog = _rt.Program(OPS_RHS, "euler", 1, 1, 3, kernels=_rt.KERNEL_TRAJECTORY)
    sim = _rt.Sim(prog)
    sim.set_solver_params(dt=1.0, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=1, max_store=2, nout=1)
    sim.set_tspan(0.0, 1.0)
    sim.set_problem(x, a)
    sim.seed_rng(1)
    sim.trajectory()
    tr = sim.get_trajectory()
    aux = tr["aux"].reshape(-1, 3, n)[0].copy()
    sim.close()
    return aux


def _bits(v):
    return np.asarray(v, dtype=np.float64).view(np.uint64)


@pytest.mark.gpu
def test_cuda_branch_free_reciprocal_and_division_are_the_ieee_operations(monkeypatch):
    """CLODE_BRANCHLESS=1 (ptx_pass.hpp rewrite_variable_divisions): 1/x and a/x of the user's right-hand side, written
    out as ptxas' own fast-path sequence with selects instead of the slow-path branch, are bit-identical to the IEEE
    operations wherever x, a and a/x are normal (hard significands included: all ones, powers of two), and for zeros,
    infinities and NaNs; subnormal divisors count as zero (the one documented deviation)."""
    monkeypatch.setenv("CLODE_BRANCHLESS", "1")
    rng = np.random.default_rng(23)
    n = 1 << 20
    x = (rng.uniform(1.0, 2.0, n) * 2.0 ** rng.integers(-300, 300, n) * rng.choice([-1.0, 1.0], n))
    a = (rng.uniform(1.0, 2.0, n) * 2.0 ** rng.integers(-300, 300, n) * rng.choice([-1.0, 1.0], n))
    # hard significands for Newton reciprocals, over many exponents; full-range exponents; IEEE specials
    k = 0
    for e in range(-1000, 1001, 5):
        for m in (0, 1, 2, (1 << 52) - 1, (1 << 52) - 2, 1 << 51, (1 << 51) - 1):
            x[k] = np.ldexp(1.0 + m * 2.0 ** -52, e)
            a[k] = 1.0 + (k % 97) / 97.0
            k += 1
    wide = slice(k, k + 100000)
    x[wide] = rng.uniform(1.0, 2.0, 100000) * 2.0 ** rng.integers(-1020, 1020, 100000)
    a[wide] = rng.uniform(1.0, 2.0, 100000) * 2.0 ** rng.integers(-960, 1020, 100000)
    sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -3.0, 1e300, 1e-300, 2.0 ** -1022, 2.0 ** 1021])
    s0 = k + 100000
    for i, xv in enumerate(sp):
        if xv == 2.0 ** -1022:
            continue  # exponent field 1: outside the corrected range, the answer is the SFU seed (documented flush zone)
        for j, av in enumerate(sp):
            x[s0 + i * sp.size + j] = xv
            a[s0 + i * sp.size + j] = av
    aux = _ops_on_gpu(x, a)
    with np.errstate(all="ignore"):
        want_r, want_q = 1.0 / x, a / x

    def check(got, want, what, tiny_dividend=None):
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), what
        normal = ~nan & (np.abs(want) >= 2.0 ** -1021) & np.isfinite(want)
        special = ~nan & ~normal & ((want == 0.0) | np.isinf(want))
        exact = normal if tiny_dividend is None else normal & ~tiny_dividend
        bad = exact & (_bits(got) != _bits(want))
        assert not bad.any(), (what, x[bad][:5], a[bad][:5], got[bad][:5], want[bad][:5])
        assert np.array_equal(_bits(got[special]), _bits(want[special])), what            # signed zeros and infinities
        rest = ~nan & ~exact & ~special                                                   # subnormal quotients, tiny dividends
        assert np.all(np.abs(got[rest] - want[rest]) <= np.maximum(2.0 ** -1074, np.abs(want[rest]) * 2.0 ** -52)), what

    check(aux[0], want_r, "reciprocal")
    check(aux[1], want_q, "division", tiny_dividend=np.abs(a) < 2.0 ** -969)


NOISE_RHS = """void getRHS(const realtype t, const realtype x_[], const realtype p_[], realtype dx_[], realtype aux_[], const realtype w_[]) {
    dx_[0] = RCONST(0.0);
    aux_[0] = w_[0];
}
"""


def _variates_on_gpu(n, rows):
    prog = _rt.Program(NOISE_RHS, "seuler", 1, 1, 1, 1, kernels=_rt.KERNEL_TRAJECTORY)
    sim = _rt.Sim(prog)
    sim.set_solver_params(dt=1.0, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=rows, max_store=rows + 1, nout=1)
    sim.set_tspan(0.0, float(rows))
    sim.set_problem(np.zeros(n), np.ones(n))
    sim.seed_rng(7)
    sim.trajectory()
    tr = sim.get_trajectory()
    keep = int(tr["n_stored"].min()) + 1
    w = tr["aux"].reshape(-1, n)[:keep].copy()
    state = sim.get_rng_state().copy()
    sim.close()
    return w, state


@pytest.mark.gpu
def test_cuda_fast_polar_variates_match_the_library_path(monkeypatch):
    """CLODE_FAST_POLAR=1 (device/fast_polar.cuh): the normal variates of the stochastic stepper are within 4 ulp of the ones
    computed as the reference writes them (library log, IEEE division and square root: 3 + 1 ulp), the RNG state words are
    identical (same draws consumed), and the sample is standard normal"""
    n, rows = 8192, 256
    monkeypatch.setenv("CLODE_FAST_POLAR", "0")
    w_lib, s_lib = _variates_on_gpu(n, rows)
    monkeypatch.setenv("CLODE_FAST_POLAR", "1")
    w_fast, s_fast = _variates_on_gpu(n, rows)
    assert np.array_equal(s_lib, s_fast)
    assert w_lib.shape == w_fast.shape and w_lib.size > 1000000
    assert np.all(np.abs(w_fast - w_lib) <= 4 * np.spacing(np.abs(w_lib)))
    assert (w_fast != w_lib).mean() < 0.9          # not a different algorithm: most variates agree to the last bit or two
    assert abs(w_fast.mean()) < 5e-3 and abs(w_fast.std() - 1.0) < 5e-3
