"""GPU parity tests proper: the NVRTC-compiled sm_100a kernels, driven through the C ABI
(libclode_rt.so via clode_b200._rt), against
  (1) the committed golden vectors from the reference's own kernels (tests/golden/),
  (2) the C restatement oracle on the same seeded inputs, and oracle/_ref where prebuilt,
  (3) size-independent properties at the BASELINE.json sizes (see test_gpu_fullsize.py).

Two tiers:
  bit-exact tier   — kernels built with bit_exact=1 (portable transcendental math, no FMA
                     contraction) vs oracle flavour "pm": EVERY output must be bit-identical,
                     i.e. identical accepted-step counts, event counts, features, final dt/t, RNG state.
  production tier  — FMA contraction + libdevice math vs oracle flavour "libm": floating-point
                     outputs within the tolerance written in each test; integer RNG state bit-identical.
"""
import os

import numpy as np
import pytest

from golden_cases import CASES, case_inputs
from oracle import ref, restate
from oracle.common import Config, Observer, Solver, seed_states
from problems import MODELS, ensemble
from util import GpuRun, assert_bit_equal, run_oracle

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------------
# (1) golden vectors from the reference kernels
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden_bit_for_bit(rt, name, golden):
    case = CASES[name]
    ns = case.get("n_store", 0)
    ts, x0, pars, sp, op, n = case_inputs(case)
    g = GpuRun(rt, case["model"], case["stepper"], case.get("observer", "basic"), ns, bit_exact=True)
    g.setup(ts, x0, pars, sp, op, seed=case.get("seed", 1))
    r = g.run(case["kind"])
    want = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/") and "/cont_" not in k}
    assert_bit_equal(r, want, name)
    if case.get("continue"):  # features(initialize=False): device-resident observer state, dt and RNG continue
        g.sim.shift_x0()
        g.sim.set_tspan(ts[1], ts[1] + (ts[1] - ts[0]))
        r2 = g.features(initialize=0)
        want2 = {k.split("/cont_", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/cont_")}
        assert_bit_equal(r2, want2, name + " (continued)")
    g.close()


# ----------------------------------------------------------------------------------------------
# (2) the oracle on the same seeded inputs, every stepper x observer, ragged sizes
STEPPERS = ["euler", "heun", "rk4", "bs23", "dopri5"]
OBSERVERS = ["basic", "basicall", "localmax", "nhood1", "nhood2", "thresh2"]


@pytest.mark.parametrize("stepper", STEPPERS)
@pytest.mark.parametrize("observer", OBSERVERS)
def test_cuda_features_bit_exact_all_pairs(rt, stepper, observer):
    model = "lorenz63" if observer in ("basic", "basicall", "localmax") else "lactotroph"
    ns = 2 if observer in ("localmax", "nhood2", "thresh2") else 0
    n = 97  # ragged: not a multiple of the warp or block size
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 4)
    fixed = stepper in ("euler", "heun", "rk4")
    sp = Solver(dt=(0.05 if model == "lactotroph" else 0.01) if fixed else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4,
                max_steps=200000)
    op = Observer(max_event_count=25, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    g = GpuRun(rt, model, stepper, observer, ns, bit_exact=True)
    g.setup(ts, x0, pars, sp, op, seed=3)
    r = g.features()
    o = run_oracle(restate.OracleLib(Config(model, stepper, observer, ns, math="pm")), "features", ts, x0, pars, sp, op, seed=3)
    assert_bit_equal(r, o, f"{model} {stepper} {observer}")
    # accepted-step counts: the kernel's own counter equals the observer's "step count" feature
    nfeat = g.sim.n_features()
    F = r["F"].reshape(nfeat, n)
    step_row = {"basic": 5, "basicall": nfeat - 1, "localmax": nfeat - 1, "nhood1": nfeat - 4, "nhood2": nfeat - 4,
                "thresh2": nfeat - 4}[observer]
    assert np.array_equal(F[step_row], r["steps"].astype(np.float64))
    g.close()


@pytest.mark.parametrize("model,stepper", [("vanderpol", "rk4"), ("lorenz63", "dopri5"), ("chay_keizer", "bs23"),
                                           ("lactotroph", "heun")])
def test_cuda_transient_and_trajectory_bit_exact(rt, model, stepper):
    n = 130
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 4)
    fixed = stepper in ("euler", "heun", "rk4")
    sp = Solver(dt=0.05 if fixed else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000, max_store=64, nout=3)
    lib = restate.OracleLib(Config(model, stepper, math="pm"))
    g = GpuRun(rt, model, stepper, bit_exact=True)
    g.setup(ts, x0, pars, sp, None, seed=9)
    assert_bit_equal(g.transient(), run_oracle(lib, "transient", ts, x0, pars, sp, None, seed=9), f"{model} transient")
    g.setup(ts, x0, pars, sp, None, seed=9)
    r, o = g.trajectory(), run_oracle(lib, "trajectory", ts, x0, pars, sp, None, seed=9)
    # rows beyond n_stored are never written by either side; compare only what was stored
    rows = o["rows"]
    nv, na = lib.n_var, lib.n_aux
    mask_t = np.arange(rows)[:, None] <= o["n_stored"][None, :]
    for k, width in (("t", 1), ("x", nv), ("dx", nv), ("aux", na)):
        if width == 0:
            continue
        a = r[k].reshape(rows, width, n)
        b = o[k].reshape(rows, width, n)
        m = np.broadcast_to(mask_t[:, None, :], a.shape)
        assert np.array_equal(a[m], b[m]), f"{model} trajectory {k}"
    assert_bit_equal(r, o, f"{model} trajectory", keys=["n_stored", "xf", "tf", "dt", "rng"])
    g.close()


def test_cuda_matches_reference_kernels_directly(rt):
    """straight against oracle/_ref (the reference sources as host C), where the .so was prebuilt"""
    cfg = Config("lactotroph", "bs23", "thresh2", 2, math="pm")
    if not os.path.exists(ref.so_path(cfg)):
        pytest.skip("oracle/_ref artefact not present on this machine")
    n = 64
    ts, x0, pars = ensemble("lactotroph", n)
    sp = Solver(dt=0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000)
    op = Observer(max_event_count=50, max_event_timestamps=2, x_up_threshold=0.3, x_down_threshold=0.2)
    g = GpuRun(rt, "lactotroph", "bs23", "thresh2", 2, bit_exact=True)
    g.setup(ts, x0, pars, sp, op, seed=7)
    assert_bit_equal(g.features(), run_oracle(ref.RefLib(cfg), "features", ts, x0, pars, sp, op, seed=7), "vs _ref")
    g.close()


# ----------------------------------------------------------------------------------------------
# stochastic runs: the per-instance RNG stream must be reproduced exactly
@pytest.mark.parametrize("bit_exact", [True, False])
def test_cuda_rng_stream_is_bit_identical(rt, bit_exact):
    n = 200
    ts, x0, pars = ensemble("lactotroph_noise", n)
    sp = Solver(dt=0.01, max_steps=1000000)
    lib = restate.OracleLib(Config("lactotroph_noise", "seuler", "basicall", math="pm" if bit_exact else "libm"))
    g = GpuRun(rt, "lactotroph_noise", "seuler", "basicall", bit_exact=bit_exact)
    g.setup((0.0, 30.0), x0, pars, sp, Observer(), seed=1)
    r, o = g.features(), run_oracle(lib, "features", (0.0, 30.0), x0, pars, sp, Observer(), seed=1)
    # integer state identical in BOTH tiers: the polar-method rejection test is computed with
    # contraction-proof arithmetic, so the number of draws never depends on the math library
    assert np.array_equal(r["rng"], o["rng"])
    want_steps = o["F"].reshape(-1, n)[-1]  # basicall: last feature row = step count
    assert np.array_equal(r["steps"].astype(np.float64), want_steps) and want_steps.min() >= 3000
    if bit_exact:
        assert_bit_equal(r, o, "seuler")
    else:
        # tolerance: Euler-Maruyama over 3001 steps with 1-ulp differences in log() and exp()
        assert np.allclose(r["xf"], o["xf"], rtol=1e-9, atol=1e-12)
        assert np.allclose(r["F"], o["F"], rtol=1e-9, atol=1e-12)
    g.close()


def test_cuda_rng_sharded_seeding_reproduces_the_unsharded_stream(rt):
    """instance i of a shard [offset, offset+m) of a global ensemble of n draws the same stream"""
    n, off, m = 96, 40, 33
    ts, x0, pars = ensemble("lactotroph_noise", n)
    sp = Solver(dt=0.01, max_steps=100000)
    full = GpuRun(rt, "lactotroph_noise", "seuler", bit_exact=True)
    full.setup((0.0, 5.0), x0, pars, sp, None, seed=11)
    rf = full.transient()
    part = GpuRun(rt, "lactotroph_noise", "seuler", bit_exact=True)
    sl = lambda a, w: a.reshape(w, n)[:, off:off + m].ravel()
    part.setup((0.0, 5.0), sl(x0, 4), sl(pars, 4), sp, None)
    part.sim.seed_rng(11, off, n)
    rp = part.transient()
    assert np.array_equal(rp["xf"], sl(rf["xf"], 4)) and np.array_equal(rp["rng"], sl(rf["rng"], 2))
    full.close(), part.close()


# ----------------------------------------------------------------------------------------------
# production tier: tolerance against the libm oracle on non-chaotic problems
def test_cuda_production_tier_tolerances(rt):
    n = 128
    # Van der Pol, rk4 (config C1 at reduced size): fixed step, 10^4 steps
    ts, x0, pars = ensemble("vanderpol", n)
    sp = Solver(dt=0.01, max_steps=1000000)
    g = GpuRun(rt, "vanderpol", "rk4", bit_exact=False)
    g.setup(ts, x0, pars, sp, None)
    r, o = g.transient(), run_oracle(restate.OracleLib(Config("vanderpol", "rk4")), "transient", ts, x0, pars, sp, None)
    assert np.array_equal(r["tf"], o["tf"]) and np.array_equal(r["steps"], np.full(n, 10000, np.uint32))
    assert np.allclose(r["xf"], o["xf"], rtol=1e-9, atol=1e-11)  # FMA vs non-FMA rounding over 10^4 steps
    g.close()
    # lactotroph, bs23 + thresh2 (config C3 at reduced size): adaptive, bursting but not chaotic
    ts, x0, pars = ensemble("lactotroph", n)
    sp = Solver(dt=0.1, dtmax=100.0, abstol=1e-6, reltol=1e-4, max_steps=10000000)
    op = Observer(max_event_count=100000, x_up_threshold=0.3, x_down_threshold=0.2)
    g = GpuRun(rt, "lactotroph", "bs23", "thresh2", bit_exact=False)
    g.setup(ts, x0, pars, sp, op)
    r = g.features()
    o = run_oracle(restate.OracleLib(Config("lactotroph", "bs23", "thresh2")), "features", ts, x0, pars, sp, op)
    nfeat = g.sim.n_features()
    F, G = r["F"].reshape(nfeat, n), o["F"].reshape(nfeat, n)
    ev = 18 + 20 + 3
    same_events = F[ev] == G[ev]
    assert same_events.mean() >= 0.95  # event counts identical for >= 95 % of instances
    # accepted-step counts within 1 %, mean period within 0.1 % where the event count agrees
    assert np.all(np.abs(F[ev + 1] - G[ev + 1]) <= 0.01 * G[ev + 1] + 2)
    k = same_events & (G[ev] > 1)
    assert np.allclose(F[2][k], G[2][k], rtol=1e-3)
    g.close()


def test_cuda_single_precision(rt):
    """single precision path (the reference's Python default): tolerance 1e-4 relative on a short fixed-step run"""
    n = 64
    ts, x0, pars = ensemble("vanderpol", n)
    sp = Solver(dt=0.01, max_steps=100000)
    g = GpuRun(rt, "vanderpol", "rk4", bit_exact=False, single=True)
    g.setup((0.0, 10.0), x0, pars, sp, None)
    r = g.transient()
    o = run_oracle(restate.OracleLib(Config("vanderpol", "rk4", single=True)), "transient", (0.0, 10.0), x0, pars, sp, None)
    assert np.allclose(r["xf"], o["xf"].astype(np.float64), rtol=1e-4, atol=1e-5)
    assert np.allclose(r["tf"], o["tf"].astype(np.float64), rtol=1e-6)
    g.close()


@pytest.mark.parametrize("model,stepper,observer", [("lorenz63", "dopri5", "basic"), ("vanderpol", "bs23", "basicall"),
                                                    ("lactotroph", "dopri5", "basicall")])
def test_cuda_single_precision_adaptive(rt, model, stepper, observer):
    """single precision, adaptive steppers: the production build forms the controller root with the SFU (lg2/ex2), the
    engine's divisions with a Newton reciprocal and the step floor from the exponent field.  Against the oracle in
    single precision (libm powf, IEEE division): accepted-step counts within 1 %, features within 5e-3 of the
    feature's range over the ensemble (measured: <= 2e-3; non-chaotic instances; FP32 round-off already moves
    individual steps, and extrema of slopes are taken at discrete step times)."""
    n = 96
    ts, x0, pars = ensemble(model, n)
    if model == "lorenz63":  # r in [0.5, 20]: fixed points, no chaos
        pars = np.concatenate([np.linspace(0.5, 20.0, n), np.full(n, 10.0), np.full(n, 8.0 / 3.0)])
    ts = (0.0, 20.0 if model != "lactotroph" else 400.0)
    sp = Solver(dt=0.01, dtmax=1.0, abstol=1e-5, reltol=1e-4, max_steps=1000000)
    g = GpuRun(rt, model, stepper, observer, bit_exact=False, single=True)
    g.setup(ts, x0, pars, sp, Observer())
    r = g.features()
    o = run_oracle(restate.OracleLib(Config(model, stepper, observer, single=True)), "features", ts, x0, pars, sp, Observer())
    nf = len(r["F"]) // n
    Fg, Fo = r["F"].reshape(nf, n), o["F"].astype(np.float64).reshape(nf, n)
    so, sg = Fo[-1], r["steps"].astype(np.float64)  # the last feature of these observers is the step count
    assert np.array_equal(Fg[-1], sg)
    assert np.all(np.abs(so - sg) <= np.maximum(3.0, 0.01 * so)), (so[:8], sg[:8])
    scale = np.maximum(np.abs(Fo).max(axis=1, keepdims=True), 1e-3)
    keep = np.ones(nf, bool)
    keep[-1] = False  # the step-count feature is compared above
    worst = (np.abs(Fg - Fo) / scale)[keep].max(axis=1)
    print("single precision, worst feature deviation / range:", worst)
    assert np.all(worst <= 5e-3), worst
    assert np.allclose(r["tf"], o["tf"].astype(np.float64), rtol=1e-5)
    g.close()


# ----------------------------------------------------------------------------------------------
# edge cases
def test_cuda_edge_sizes_and_limits(rt):
    lib = restate.OracleLib(Config("lorenz63", "dopri5", "basic", math="pm"))
    sp = Solver(dt=0.1, dtmax=1.0, abstol=1e-6, reltol=1e-5, max_steps=37)  # max_steps cut-off mid-run
    for n in (1, 31, 32, 33, 65):
        ts, x0, pars = ensemble("lorenz63", n)
        g = GpuRun(rt, "lorenz63", "dopri5", "basic", bit_exact=True)
        g.setup(ts, x0, pars, sp, Observer())
        r = g.features()
        assert_bit_equal(r, run_oracle(lib, "features", ts, x0, pars, sp, Observer()), f"n={n}")
        assert np.array_equal(r["steps"], np.full(n, 37, np.uint32))
        g.close()


def test_cuda_zero_length_interval_and_nan_inputs(rt):
    n = 16
    ts, x0, pars = ensemble("lorenz63", n)
    lib = restate.OracleLib(Config("lorenz63", "dopri5", "basicall", math="pm"))
    sp = Solver(dt=0.1, dtmax=1.0, reltol=1e-5, max_steps=500)
    x0 = x0.copy()
    x0[3] = np.nan  # a NaN state: the controller's fmax-based norm accepts every step
    for tspan in ((0.0, 0.0), (0.0, 2.0)):
        g = GpuRun(rt, "lorenz63", "dopri5", "basicall", bit_exact=True)
        g.setup(tspan, x0, pars, sp, Observer())
        assert_bit_equal(g.features(), run_oracle(lib, "features", tspan, x0, pars, sp, Observer()), f"tspan={tspan}")
        g.close()


def test_cuda_abi_error_behaviour(rt):
    from problems import rhs_source
    sim = rt.Sim(rt.Program(rhs_source("lorenz63"), "rk4", 3, 3, 1, kernels=rt.KERNEL_TRANSIENT))
    with pytest.raises(rt.RtError):  # no problem data yet
        sim.transient()
    with pytest.raises(rt.RtError):  # features kernels were not built
        sim.features()
    sim.set_solver_params(dt=0.01)
    with pytest.raises(ValueError):
        sim.set_problem(np.ones(7), np.ones(9))
    sim.set_problem(np.ones(6), np.ones(6))
    with pytest.raises(rt.RtError):  # wrong-sized x0
        sim.set_x0(np.ones(5))
    sim.close()


# ----------------------------------------------------------------------------------------------
# persistent-thread work queue with per-lane refill: results must not depend on which lane ran what
@pytest.mark.parametrize("model,stepper,observer", [("lorenz63", "dopri5", "basic"), ("lactotroph", "bs23", "thresh2")])
def test_cuda_block_order_does_not_change_results(rt, model, stepper, observer, monkeypatch):
    """the block -> chunk mapping (forward / reverse / decided from the previous launch's step counts) is a schedule:
    every output is bit-identical, on the first call (no history) and on a repeated call (history from the first)"""
    n = 3000 + 7  # ragged: the last block is partial
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 8)
    sp = Solver(dt=0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000)
    op = Observer(max_event_count=8, x_up_threshold=0.3, x_down_threshold=0.2)
    results = {}
    for order in ("forward", "reverse", "auto"):
        monkeypatch.setenv("CLODE_BLOCK_ORDER", order)
        g = GpuRun(rt, model, stepper, observer, bit_exact=True)
        g.setup(ts, x0, pars, sp, op, seed=2)
        first = g.features()
        g.setup(ts, x0, pars, sp, op, seed=2)   # same ensemble again: `auto` now has the first call's step counts
        second = g.features()
        g.close()
        assert_bit_equal(second, first, f"{order}: repeated call")
        results[order] = first
    assert_bit_equal(results["reverse"], results["forward"], "reverse vs forward")
    assert_bit_equal(results["auto"], results["forward"], "auto vs forward")


@pytest.mark.parametrize("model,stepper,observer,kind", [
    ("lorenz63", "dopri5", "localmax", "features"),
    ("lactotroph", "bs23", "thresh2", "features"),   # two-pass: the warm-up kernel also runs from the queue
    ("lorenz63", "dopri5", "basic", "transient"),
    ("chay_keizer", "dopri5", "basic", "trajectory"),
    ("lactotroph_noise", "seuler", "basicall", "features"),
])
def test_cuda_work_queue_lane_refill_bit_exact(rt, model, stepper, observer, kind):
    n = 5000 + 13  # many refill rounds per warp, ragged tail
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 8)
    rng_perm = np.random.default_rng(5).permutation(n)  # shuffle: heterogeneous step counts inside every warp
    nv = len(x0) // n
    npar = len(pars) // n
    x0 = x0.reshape(nv, n)[:, rng_perm].ravel()
    pars = pars.reshape(npar, n)[:, rng_perm].ravel()
    ns = 2 if observer in ("localmax", "thresh2") else 0
    fixed = stepper == "seuler"
    sp = Solver(dt=0.01 if fixed else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000, max_store=40, nout=3)
    op = Observer(max_event_count=8, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2)
    plain = GpuRun(rt, model, stepper, observer, ns, bit_exact=True, work_queue=False)
    queue = GpuRun(rt, model, stepper, observer, ns, bit_exact=True, work_queue=True)
    out = []
    for g in (plain, queue):
        g.setup(ts, x0, pars, sp, op if kind == "features" else None, seed=2)
        out.append(g.run(kind))
        g.close()
    a, b = out
    if kind == "trajectory":
        rows = a["rows"]
        stored = np.arange(rows)[:, None] <= a["n_stored"][None, :]
        for k, width in (("t", 1), ("x", nv), ("dx", nv)):
            m = np.broadcast_to(stored[:, None, :], (rows, width, n))
            assert np.array_equal(a[k].reshape(rows, width, n)[m], b[k].reshape(rows, width, n)[m]), k
        assert_bit_equal(b, a, "queue vs plain", keys=["n_stored", "xf", "tf", "dt", "rng", "steps"])
    else:
        assert_bit_equal(b, a, "queue vs plain")
    assert a["steps"].max() > a["steps"].min() or stepper == "seuler"  # step counts differ between lanes
    # and both equal the oracle on a sample of instances
    sub = np.arange(0, n, 97)
    lib = restate.OracleLib(Config(model, stepper, observer if kind == "features" else "basic", ns if kind == "features" else 0, math="pm"))
    pick = lambda v, w: np.asarray(v).reshape(w, n)[:, sub].ravel()
    dt, rng = np.full(sub.size, sp.dt), pick(seed_states(2, n), 2)
    if kind == "features":
        o = lib.features(ts, pick(x0, nv), pick(pars, npar), sp, op, dt, rng)
        assert np.array_equal(pick(b["F"], len(b["F"]) // n), o["F"])
    elif kind == "transient":
        o = lib.transient(ts, pick(x0, nv), pick(pars, npar), sp, dt, rng)
    else:
        o = lib.trajectory(ts, pick(x0, nv), pick(pars, npar), sp, dt, rng)
        assert np.array_equal(b["n_stored"][sub], o["n_stored"])
    assert np.array_equal(pick(b["xf"], nv), o["xf"]) and np.array_equal(pick(b["rng"], 2), o["rng"])


# ----------------------------------------------------------------------------------------------
# trajectory rows staged in shared memory and written with TMA bulk copies (fixed-step methods)
@pytest.mark.parametrize("model,stepper,n,nout,max_store", [
    ("chay_keizer", "rk4", 1000, 1, 50),      # nout = 1, stops on max_store
    ("chay_keizer", "euler", 258, 3, 400),    # last block partially filled, stops on t_end
    ("thompson_a1", "heun", 131, 2, 60),      # odd n: alignment fallback to per-thread stores
    ("lactotroph_noise", "seuler", 512, 4, 30),
])
def test_cuda_staged_trajectory_bit_exact(rt, model, stepper, n, nout, max_store):
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[0] + (ts[1] - ts[0]) / 10)
    sp = Solver(dt=0.05, dtmax=1.0, max_steps=100000, max_store=max_store, nout=nout)
    lib = restate.OracleLib(Config(model, stepper, math="pm"))
    o = run_oracle(lib, "trajectory", ts, x0, pars, sp, None, seed=4)
    g = GpuRun(rt, model, stepper, bit_exact=True, staged=True)
    g.setup(ts, x0, pars, sp, None, seed=4)
    r = g.trajectory()
    rows, nv, na = o["rows"], lib.n_var, lib.n_aux
    stored = np.arange(rows)[:, None] <= o["n_stored"][None, :]
    for k, width in (("t", 1), ("x", nv), ("dx", nv), ("aux", na)):
        if width:
            m = np.broadcast_to(stored[:, None, :], (rows, width, n))
            assert np.array_equal(r[k].reshape(rows, width, n)[m], o[k].reshape(rows, width, n)[m]), k
    assert_bit_equal(r, o, "staged trajectory", keys=["n_stored", "xf", "tf", "dt", "rng"])
    g.close()


# ----------------------------------------------------------------------------------------------
# streamed trajectory: chunked launches, copy-out overlapped with the next chunk
@pytest.mark.parametrize("model,stepper,n,nout,max_store,chunk_rows,pinned,bit_exact", [
    ("chay_keizer", "rk4", 1000, 1, 50, 7, True, True),          # stops on max_store; last chunk partial
    ("chay_keizer", "rk4", 1000, 1, 50, 1000, False, True),      # one chunk holds everything
    ("lorenz63", "dopri5", 777, 1, 400, 64, True, True),         # adaptive: instances finish in different chunks
    ("lorenz63", "dopri5", 777, 2, 400, 33, False, False),       # production arithmetic, pageable host arrays
    ("lactotroph", "bs23", 300, 1, 120, 16, True, True),
    ("lactotroph_noise", "seuler", 512, 4, 30, 4, True, True),   # RNG state, cached variate, pending noise
    ("vanderpol", "heun", 131, 3, 90, 1, False, True),           # one row per launch
])
def test_cuda_streamed_trajectory_equals_single_launch(rt, model, stepper, n, nout, max_store, chunk_rows, pinned, bit_exact):
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[0] + (ts[1] - ts[0]) / 10)
    sp = Solver(dt=0.05, dtmax=1.0, abstol=1e-6, reltol=1e-5, max_steps=100000, max_store=max_store, nout=nout)
    g = GpuRun(rt, model, stepper, bit_exact=bit_exact)
    g.setup(ts, x0, pars, sp, None, seed=4)
    whole = g.trajectory()
    launches = g.sim.launch_count()
    g.setup(ts, x0, pars, sp, None, seed=4)
    g.sim.set_dt(np.full(n, sp.dt))
    parts = dict(g.sim.trajectory_stream(chunk_rows, pinned=pinned))
    parts.update(g.common())
    assert g.sim.launch_count() - launches <= -(-(max_store + 1) // min(chunk_rows, max_store + 1))
    if bit_exact:  # and against the oracle
        o = run_oracle(restate.OracleLib(Config(model, stepper, math="pm")), "trajectory", ts, x0, pars, sp, None, seed=4)
        assert_bit_equal(whole, o, "single launch", keys=["n_stored", "xf", "tf", "dt", "rng"])
    nv, na = MODELS[model][0], MODELS[model][2]
    rows = max_store
    # the FSAL slope is recomputed at a chunk boundary; with FMA contraction (production build) the two inlined
    # copies of the user's RHS may round differently, so only the bit-exact tier is compared bit for bit
    same = (lambda a, b: np.array_equal(a, b)) if bit_exact else (lambda a, b: np.allclose(a, b, rtol=1e-6, atol=1e-9))
    assert np.array_equal(whole["n_stored"], parts["n_stored"]) or not bit_exact
    kept = np.minimum(parts["n_stored"].astype(np.int64) + 1, rows)
    for k, width in (("t", 1), ("x", nv), ("dx", nv), ("aux", na)):
        if not width:
            continue
        a = whole[k][:rows * width * n].reshape(rows, width, n)
        b = np.asarray(parts[k]).reshape(rows, width, n)
        live = np.broadcast_to(np.arange(rows)[:, None, None] < kept[None, None, :], a.shape)
        if bit_exact:
            assert np.array_equal(a[live], b[live]), k
        else:  # chaotic instances decorrelate: compare the early part of every trajectory
            early = live & (np.arange(rows)[:, None, None] < 40)
            assert np.allclose(a[early], b[early], rtol=1e-5, atol=1e-7), k
        assert not b[~live].any(), f"{k}: rows beyond n_stored must read zero"
    for k in ("xf", "tf", "dt", "rng", "steps"):
        if bit_exact:
            assert np.array_equal(whole[k], parts[k]), k
    g.close()


def test_cuda_streamed_trajectory_single_precision_and_skipped_outputs(rt):
    n, max_store = 200, 60
    ts, x0, pars = ensemble("chay_keizer", n)
    sp = Solver(dt=0.5, dtmax=1.0, max_steps=100000, max_store=max_store, nout=1)
    g = GpuRun(rt, "chay_keizer", "rk4", bit_exact=False, single=True)
    g.setup((0.0, 100.0), x0, pars, sp, None, seed=1)
    whole = g.trajectory()
    g.setup((0.0, 100.0), x0, pars, sp, None, seed=1)
    g.sim.set_dt(np.full(n, sp.dt))
    parts = g.sim.trajectory_stream(9, want=("t", "x"))
    assert parts["dx"] is None
    assert np.array_equal(parts["n_stored"], whole["n_stored"])
    rows = max_store
    assert np.array_equal(whole["t"][:rows * n], parts["t"]) and np.array_equal(whole["x"][:rows * n * 3], parts["x"])
    g.close()


# ----------------------------------------------------------------------------------------------
# observer state kept in a shared-memory slot (out-of-line observer step) instead of registers
@pytest.mark.parametrize("model,stepper,observer", [("lactotroph", "bs23", "thresh2"), ("lactotroph", "dopri5", "nhood2"),
                                                    ("lorenz63", "dopri5", "localmax"), ("lactotroph", "rk4", "nhood1"),
                                                    ("lorenz63", "bs23", "basicall")])
def test_cuda_observer_in_shared_memory_bit_exact(rt, model, stepper, observer):
    n = 333
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 4)
    ns = 2 if observer in ("localmax", "nhood2", "thresh2") else 0
    sp = Solver(dt=0.05 if stepper == "rk4" else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000)
    op = Observer(max_event_count=25, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    g = GpuRun(rt, model, stepper, observer, ns, bit_exact=True, obs_smem=True)
    g.setup(ts, x0, pars, sp, op, seed=3)
    r = g.features()
    o = run_oracle(restate.OracleLib(Config(model, stepper, observer, ns, math="pm")), "features", ts, x0, pars, sp, op, seed=3)
    assert_bit_equal(r, o, f"{model} {stepper} {observer} (observer in shared memory)")
    # continuation: the slot is re-loaded from the SoA record at the start of the next call
    g.sim.shift_x0()
    g.sim.set_tspan(ts[1], ts[1] + (ts[1] - ts[0]))
    r2 = g.features(initialize=0)
    lib = restate.OracleLib(Config(model, stepper, observer, ns, math="pm"))
    o1 = run_oracle(lib, "features", ts, x0, pars, sp, op, seed=3)
    o2 = lib.features((ts[1], ts[1] + (ts[1] - ts[0])), o1["xf"], pars, sp, op, o1["dt"], o1["rng"], initialize=False)
    assert_bit_equal(r2, o2, "continued")
    g.close()


@pytest.mark.parametrize("model,stepper,observer", [("lactotroph", "bs23", "thresh2"), ("lactotroph", "dopri5", "nhood2"),
                                                    ("lorenz63", "dopri5", "localmax"), ("lactotroph", "rk4", "nhood1"),
                                                    ("lorenz63", "bs23", "basicall")])
def test_cuda_extents_in_shared_memory_bit_exact(rt, model, stepper, observer, monkeypatch):
    """CLODE_EXT_SMEM=1: the per-variable extents / means of the multi-variable observers live in shared memory
    (observers.cuh); placement only — every output is bit-identical to the oracle, also across a continued call"""
    monkeypatch.setenv("CLODE_EXT_SMEM", "1")
    n = 333
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 4)
    ns = 2 if observer in ("localmax", "nhood2", "thresh2") else 0
    sp = Solver(dt=0.05 if stepper == "rk4" else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000)
    op = Observer(max_event_count=25, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    g = GpuRun(rt, model, stepper, observer, ns, bit_exact=True)
    assert "-DCLODE_EXT_SMEM" in rt.program_source(g.prog)
    g.setup(ts, x0, pars, sp, op, seed=3)
    r = g.features()
    lib = restate.OracleLib(Config(model, stepper, observer, ns, math="pm"))
    o = run_oracle(lib, "features", ts, x0, pars, sp, op, seed=3)
    assert_bit_equal(r, o, f"{model} {stepper} {observer} (extents in shared memory)")
    g.sim.shift_x0()
    g.sim.set_tspan(ts[1], ts[1] + (ts[1] - ts[0]))
    r2 = g.features(initialize=0)
    o2 = lib.features((ts[1], ts[1] + (ts[1] - ts[0])), o["xf"], pars, sp, op, o["dt"], o["rng"], initialize=False)
    assert_bit_equal(r2, o2, "continued")
    g.close()
