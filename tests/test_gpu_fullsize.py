"""Parity at the BASELINE.json sizes.  The oracle cannot integrate 10^6 instances in seconds, but
instances are independent, so a size-independent property pins the full-size run: the results of ANY
subset of instances of the big launch must be bit-identical (bit-exact tier) to the oracle run on just
that subset.  Subsets are drawn at random over the whole index range, so a layout / indexing / sharding
error anywhere in the 2^20-wide arrays shows up.  Further properties: step-count checksum equality
between the kernel's counter and the feature column, RNG state words of untouched instances, and
trajectory rows read straight from the 29 GB device buffers."""
import numpy as np
import pytest

from oracle import restate
from oracle.common import MODELS, Config, Observer, Solver
from clode_b200 import sharding
from problems import rhs_source

pytestmark = pytest.mark.gpu


def _grid(workload, n):
    import bench
    w = bench.workload(workload, n, np.arange(n))
    return w


def _sim(rt, w, kernels, bit_exact=True, n_store=0):
    nv, npar, na, nw = MODELS[w["model"]]
    prog = rt.Program(rhs_source(w["model"]), w["stepper"], nv, npar, na, nw, observer=w["observer"], kernels=kernels,
                      bit_exact=bit_exact, n_store_events=n_store)
    sim = rt.Sim(prog)
    sim.set_solver_params(**w["solver"])
    sim.set_observer_params(**w["observer_params"])
    sim.set_tspan(*w["tspan"])
    return sim, (nv, npar, na, nw)


def _subset(n, k=384, seed=0):
    rng = np.random.default_rng(seed)
    return np.sort(np.concatenate([[0, 1, 31, 32, n - 1], rng.choice(n, k, replace=False)]))


def test_c2_lorenz_dopri5_features_2e20(rt):
    n = 1 << 20
    w = _grid("C2", n)
    sim, (nv, npar, na, nw) = _sim(rt, w, rt.KERNEL_FEATURES)
    sim.set_problem(w["x0"], w["pars"])
    sim.seed_rng(1)
    sim.features(1)
    F, xf, tf, dt, steps = sim.get_f().reshape(6, n), sim.get_xf().reshape(nv, n), sim.get_tf(), sim.get_dt(), sim.get_steps()
    # checksum property: the kernel's accepted-step counter equals the "step count" feature everywhere
    assert np.array_equal(F[5], steps.astype(np.float64)) and int(steps.astype(np.int64).sum()) > 4e9
    assert np.all(tf >= 100.0) and np.all(np.isfinite(F))
    sub = _subset(n)
    lib = restate.OracleLib(Config("lorenz63", "dopri5", "basic", math="pm"))
    sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
    o = lib.features(w["tspan"], sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub), sp, op,
                     np.full(sub.size, sp.dt), sharding.seed_states_for(1, n, sub), nthreads=8)
    assert np.array_equal(F[:, sub].ravel(), o["F"])          # incl. identical accepted-step counts
    assert np.array_equal(xf[:, sub].ravel(), o["xf"]) and np.array_equal(tf[sub], o["tf"]) and np.array_equal(dt[sub], o["dt"])
    sim.close()


def test_c3_lactotroph_thresh2_bs23_1024x1024(rt):
    n = 1 << 20
    w = _grid("C3", n)
    w["tspan"] = (0.0, 2000.0)  # a fifth of the benchmark horizon keeps the two-pass bit-exact build to seconds
    sim, (nv, npar, na, nw) = _sim(rt, w, rt.KERNEL_FEATURES)
    sim.set_problem(w["x0"], w["pars"])
    sim.seed_rng(1)
    sim.features(1)
    nf = sim.n_features()
    F, xf = sim.get_f().reshape(nf, n), sim.get_xf().reshape(nv, n)
    ev, st = 18 + 5 * nv + 3 * na, 18 + 5 * nv + 3 * na + 1
    assert np.array_equal(F[st], sim.get_steps().astype(np.float64))
    sub = _subset(n, 256)
    lib = restate.OracleLib(Config("lactotroph", "bs23", "thresh2", math="pm"))
    sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
    o = lib.features(w["tspan"], sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub), sp, op,
                     np.full(sub.size, sp.dt), sharding.seed_states_for(1, n, sub), nthreads=8)
    G = o["F"].reshape(nf, sub.size)
    assert np.array_equal(F[ev, sub], G[ev]) and np.array_equal(F[st, sub], G[st])  # event and step counts
    assert np.array_equal(F[:, sub], G) and np.array_equal(xf[:, sub].ravel(), o["xf"])
    assert F[ev].max() > 10  # the grid really contains bursting / spiking cells
    sim.close()


def test_c4_stochastic_rng_streams_4m(rt):
    n = 1 << 22
    w = _grid("C4", n)
    w["tspan"] = (0.0, 10.0)  # 1001 steps per instance: 4.2e9 RNG-driven steps in the launch
    sim, (nv, npar, na, nw) = _sim(rt, w, rt.KERNEL_FEATURES)
    sim.set_problem(w["x0"], w["pars"])
    sim.seed_rng(1)  # CLODE::seedRNG(1): instance i gets {1 + i, 1 + n + i}
    sim.features(1)
    rng_state = sim.get_rng_state().reshape(2, n)
    nf = sim.n_features()
    F = sim.get_f().reshape(nf, n)
    sub = _subset(n, 256)
    lib = restate.OracleLib(Config("lactotroph_noise", "seuler", "basicall", math="pm"))
    sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
    o = lib.features(w["tspan"], sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub), sp, op,
                     np.full(sub.size, sp.dt), sharding.seed_states_for(1, n, sub), nthreads=8)
    assert np.array_equal(rng_state[:, sub].ravel(), o["rng"])   # per-instance streams reproduced bit for bit
    assert np.array_equal(F[:, sub].ravel(), o["F"])
    # identical parameters, different noise: the ensemble must actually be spread out
    assert np.unique(F[0]).size > n // 2
    sim.close()


def test_c5_trajectory_2000_points_x_256k(rt):
    import torch

    n = 1 << 18
    w = _grid("C5", n)
    sim, (nv, npar, na, nw) = _sim(rt, w, rt.KERNEL_TRAJECTORY)
    sim.set_problem(w["x0"], w["pars"])
    sim.seed_rng(1)
    sim.trajectory()
    rows = w["solver"]["max_store"] + 1
    n_stored = sim.get_trajectory_counts()
    assert np.array_equal(n_stored, np.full(n, 2000))
    sub = _subset(n, 64)
    lib = restate.OracleLib(Config("chay_keizer", "rk4", math="pm"))
    sp = Solver(**w["solver"])
    o = lib.trajectory(w["tspan"], sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub), sp,
                       np.full(sub.size, sp.dt), sharding.seed_states_for(1, n, sub), nthreads=8)
    # read the sampled columns straight out of the 29 GB device buffers (zero-copy view, gather on the GPU)
    class Dev:
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    idx = torch.as_tensor(sub, device="cuda")
    for which, width, key in ((rt.BUF_T, 1, "t"), (rt.BUF_X, nv, "x"), (rt.BUF_DX, nv, "dx")):
        ptr, nbytes, _ = sim.device_buffer(which)
        assert nbytes == 8 * rows * width * n
        dev = torch.as_tensor(Dev(ptr, rows * width * n), device="cuda").view(rows * width, n)
        got = dev[:, idx].cpu().numpy()
        assert np.array_equal(got[: 2001 * width].ravel(), o[key].reshape(rows * width, sub.size)[: 2001 * width].ravel()), key
    assert np.array_equal(sim.get_xf().reshape(nv, n)[:, sub].ravel(), o["xf"])
    sim.close()


def test_c1_van_der_pol_rk4_transient_4096(rt):
    """BASELINE config 1 in full: Van der Pol, transient, rk4 dt = 0.01, t in [0, 100], 4096-point mu grid, double —
    the oracle integrates all of it, so the comparison is exhaustive (bit-exact tier) and toleranced (production)."""
    from problems import ensemble
    n = 4096
    ts, x0, pars = ensemble("vanderpol", n)
    sp = Solver(dt=0.01, max_steps=1000000)
    lib_pm = restate.OracleLib(Config("vanderpol", "rk4", math="pm"))
    o = lib_pm.transient(ts, x0, pars, sp, np.full(n, sp.dt), sharding.seed_states(1, n, 0, n), nthreads=8)
    for bit_exact in (True, False):
        sim = rt.Sim(rt.Program(rhs_source("vanderpol"), "rk4", 2, 1, 0, 0, kernels=rt.KERNEL_TRANSIENT, bit_exact=bit_exact))
        sim.set_solver_params(dt=0.01, max_steps=1000000)
        sim.set_tspan(*ts)
        sim.set_problem(x0, pars)
        sim.seed_rng(1)
        sim.transient()
        xf, tf, steps = sim.get_xf(), sim.get_tf(), sim.get_steps()
        assert np.array_equal(tf, o["tf"]) and np.all(steps == steps[0]) and steps[0] in (10000, 10001)
        if bit_exact:
            assert np.array_equal(xf, o["xf"])
        else:
            assert np.allclose(xf, o["xf"], rtol=1e-8, atol=1e-10)  # FMA contraction over 10^4 steps, stiff for mu ~ 10
        sim.close()
