"""The product's CUDA device code (clode_b200/csrc/device/*.cuh), compiled with g++ behind
tests/emu/cuda_emu.h and run on the CPU, against the oracle and the golden vectors — bit for bit.
This is how kernel logic is regression-tested in the GPU-less build container; the GPU suite
(tests/test_gpu_parity.py) repeats the same comparisons with the real NVRTC-compiled kernels.
CPU only."""
import numpy as np
import pytest

from emu import EmuLib
from golden_cases import CASES, case_inputs
from oracle import restate
from oracle.common import Config, Observer, Solver, seed_states
from problems import ensemble
from util import assert_bit_equal, run_oracle


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_code_matches_golden(name, golden):
    case = CASES[name]
    cfg = Config(case["model"], case["stepper"], case.get("observer", "basic"), case.get("n_store", 0), math="pm")
    lib = EmuLib(cfg)
    ts, x0, pars, sp, op, n = case_inputs(case)
    r = run_oracle(lib, case["kind"], ts, x0, pars, sp, op, seed=case.get("seed", 1))
    want = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/") and "/cont_" not in k}
    assert_bit_equal(r, want, name)
    if case.get("continue"):
        ts2 = (ts[1], ts[1] + (ts[1] - ts[0]))
        r2 = lib.features(ts2, r["xf"], pars, sp, op, r["dt"], r["rng"], initialize=False)
        want2 = {k.split("/cont_", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/cont_")}
        assert_bit_equal(r2, want2, name + " (continued)")


@pytest.mark.parametrize("model,stepper", [("lorenz63", "dopri5"), ("lorenz63", "rk4"), ("lactotroph", "bs23"),
                                           ("vanderpol", "heun"), ("sine_drive", "euler")])
@pytest.mark.parametrize("observer", ["basic", "basicall", "localmax", "nhood1", "nhood2", "thresh2"])
def test_device_code_matches_oracle(model, stepper, observer):
    ns = 2 if observer in ("localmax", "nhood2", "thresh2") else 0
    n = 5
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 5)
    fixed = stepper in ("euler", "heun", "rk4")
    sp = Solver(dt=0.01 if fixed else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=100000, max_store=90, nout=7)
    op = Observer(max_event_count=30, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    for math in ("pm", "libm"):
        cfg = Config(model, stepper, observer, ns, math=math)
        A, B = EmuLib(cfg), restate.OracleLib(cfg)
        assert_bit_equal(run_oracle(A, "features", ts, x0, pars, sp, op), run_oracle(B, "features", ts, x0, pars, sp, op),
                         cfg.tag)


def test_device_code_feature_and_event_variable_indices():
    """fVarIx / eVarIx are baked into the device code as constants; the oracle takes them at run time"""
    cfg = Config("lactotroph", "dopri5", "thresh2", 2, math="pm")
    n = 5
    ts, x0, pars = ensemble("lactotroph", n)
    sp = Solver(dt=0.1, dtmax=10.0, reltol=1e-4, max_steps=100000)
    for e_var, f_var in [(0, 3), (3, 0), (1, 2)]:
        op = Observer(e_var_ix=e_var, f_var_ix=f_var, max_event_count=30, max_event_timestamps=2,
                      x_up_threshold=0.3, x_down_threshold=0.2)
        A, B = EmuLib(cfg, f_var_ix=f_var, e_var_ix=e_var), restate.OracleLib(cfg)
        assert_bit_equal(run_oracle(A, "features", (0.0, 1000.0), x0, pars, sp, op),
                         run_oracle(B, "features", (0.0, 1000.0), x0, pars, sp, op), f"eVar={e_var} fVar={f_var}")


def test_device_code_single_precision():
    cfg = Config("vanderpol", "dopri5", "thresh2", single=True)
    n = 6
    ts, x0, pars = ensemble("vanderpol", n)
    sp = Solver(dt=0.1, dtmax=1.0, reltol=1e-3, max_steps=20000)
    op = Observer(x_up_threshold=0.3, x_down_threshold=0.2)
    A, B = EmuLib(cfg), restate.OracleLib(cfg)
    assert_bit_equal(run_oracle(A, "features", ts, x0, pars, sp, op), run_oracle(B, "features", ts, x0, pars, sp, op),
                     "single precision")


def _streamed_equals_whole(whole, parts, n, widths, label):
    """rows < n_stored (capped at the max_store rows the API returns) of every instance are bit-identical; rows
    beyond read zero; final state, RNG words and counters are identical"""
    rows = parts["rows"]
    assert whole["rows"] == rows + 1
    for key in ("n_stored", "xf", "tf", "dt", "rng", "steps"):
        assert np.array_equal(np.asarray(whole[key]).view(np.uint8), np.asarray(parts[key]).view(np.uint8)), f"{label}: {key}"
    kept = np.minimum(whole["n_stored"].astype(np.int64) + 1, rows)  # row index n_stored is the last one written
    for key, width in widths.items():
        if width == 0:
            continue
        a = whole[key][:rows * width * n].reshape(rows, width, n)
        b = parts[key].reshape(rows, width, n)
        live = np.arange(rows)[:, None, None] < kept[None, None, :]
        live = np.broadcast_to(live, a.shape)
        assert np.array_equal(a[live].view(np.uint64), b[live].view(np.uint64)), f"{label}: {key}"
        assert not b[~live].any(), f"{label}: {key} rows beyond n_stored must read zero"


@pytest.mark.parametrize("model,stepper,dt,nout,max_store,chunks", [
    ("chay_keizer", "rk4", 0.5, 1, 120, (1, 7, 50, 121, 500)),       # fixed step, cut off by max_store
    ("vanderpol", "heun", 0.05, 3, 400, (16, 64)),                    # nout > 1, runs to t_end
    ("lorenz63", "dopri5", 0.01, 1, 300, (5, 33, 128)),               # adaptive: instances store at different rates
    ("lactotroph", "bs23", 0.1, 2, 150, (10, 64)),
    ("lactotroph_noise", "seuler", 0.05, 1, 90, (1, 8, 40)),          # RNG words, cached variate, pending noise
])
def test_streamed_trajectory_matches_single_launch(model, stepper, dt, nout, max_store, chunks):
    """the chunked trajectory (clode_sim_trajectory_stream's launch sequence, replayed on the
    emulated device code) produces the rows, counts and final state of the single-launch kernel bit for bit."""
    n = 6
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[0] + (ts[1] - ts[0]) / 20)
    sp = Solver(dt=dt, dtmax=1.0, abstol=1e-6, reltol=1e-5, max_steps=100000, max_store=max_store, nout=nout)
    cfg = Config(model, stepper, math="pm")
    lib = EmuLib(cfg)
    whole = run_oracle(lib, "trajectory", ts, x0, pars, sp, None)
    widths = dict(t=1, x=lib.n_var, dx=lib.n_var, aux=lib.n_aux)
    for chunk_rows in chunks:
        d, rng = np.full(n, sp.dt), __import__("oracle.common", fromlist=["seed_states"]).seed_states(1, n)
        parts = lib.trajectory_stream(ts, x0, pars, sp, d, rng, chunk_rows)
        _streamed_equals_whole(whole, parts, n, widths, f"{model}/{stepper} chunk_rows={chunk_rows}")
        assert parts["launches"] <= -(-(max_store + 1) // min(chunk_rows, max_store + 1))


@pytest.mark.parametrize("order", [1])
def test_block_order_only_permutes_the_schedule(order):
    """KernelArgs::block_order = 1 walks the ensemble backwards: every instance is still integrated exactly once, so all outputs are unchanged (ragged n: 9 blocks, last one partial)"""
    cfg = Config("lorenz63", "dopri5", "basic", math="pm")
    lib = EmuLib(cfg)
    n = 8 * 128 + 37
    ts, x0, pars = ensemble("lorenz63", n)
    sp = Solver(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, max_steps=100000)
    dt, rng = np.full(n, sp.dt), seed_states(1, n)
    want = lib.features((0.0, 2.0), x0, pars, sp, Observer(), dt, rng)
    lib.block_order = order
    got = lib.features((0.0, 2.0), x0, pars, sp, Observer(), dt, rng)
    lib.block_order = 0
    assert_bit_equal(got, want, f"block_order={order}")


@pytest.mark.parametrize("model, stepper, observer, n_store, tspan", [
    ("lorenz63", "dopri5", "basic", 0, (0.0, 6.0)),
    ("lorenz63", "dopri5", "localmax", 2, (0.0, 6.0)),
    ("lorenz63", "bs23", "nhood1", 0, (0.0, 6.0)),
    ("lorenz63", "bs23", "nhood2", 2, (0.0, 6.0)),       # two-pass: the warm-up pass is chunked as well
    ("lactotroph", "bs23", "thresh2", 2, (0.0, 600.0)),  # C3's kernel pair
    ("vanderpol", "dopri5", "thresh2", 0, (0.0, 30.0)),
    ("lactotroph", "dopri5", "basicall", 0, (0.0, 300.0)),
])
@pytest.mark.parametrize("math, contract", [("pm", "off"), ("libm", "fast")])
def test_scheduled_time_loop_is_bit_identical_to_one_launch(model, stepper, observer, n_store, tspan, math, contract):
    """Cost-sorted chunked execution (kernels.cuh "Scheduling", clode_rt.cpp run_loop): a time loop cut into launches
    of bounded attempt budget, with the unfinished instances re-ordered arbitrarily in between, parks and resumes every
    instance without changing a bit of its results — features, accepted-step counts, final state, dt — in the bit-exact
    arithmetic and under FMA contraction alike (the FSAL slope is stored, never recomputed)."""
    cfg = Config(model, stepper, observer, n_store, math=math, contract=contract)
    lib = EmuLib(cfg)
    n = 2 * 128 + 19
    _, x0, pars = ensemble(model, n)
    sp = Solver(dt=0.01, dtmax=1.0 if model != "lactotroph" else 50.0, abstol=1e-6, reltol=1e-5, max_steps=100000)
    op = Observer(max_event_count=50, max_event_timestamps=n_store, x_up_threshold=0.3, x_down_threshold=0.2)
    dt, rng = np.full(n, sp.dt), seed_states(1, n)
    want = lib.features(tspan, x0, pars, sp, op, dt, rng)
    for budgets, seed in (((1,), 0), ((7, 50, 300), 1), ((64, 64, 64, 64, 64, 64), 2)):
        got = lib.features_chunked(tspan, x0, pars, sp, op, dt, rng, budgets=budgets, seed=seed)
        assert got.pop("launches") >= 2
        assert_bit_equal(got, want, f"{model}/{stepper}/{observer} {math} budgets={budgets}")
    want_t = lib.transient(tspan, x0, pars, sp, dt, rng)
    got_t = lib.transient_chunked(tspan, x0, pars, sp, dt, rng, budgets=(5, 40, 200), seed=3)
    got_t.pop("launches")
    assert_bit_equal(got_t, want_t, f"{model}/{stepper} transient {math}")


def test_scheduled_time_loop_handles_max_steps_and_empty_rounds():
    """an instance that stops on max_steps (not on t_end) finishes inside a round like any other; rounds that find
    nothing left to do are harmless"""
    cfg = Config("lorenz63", "dopri5", "basic", math="pm")
    lib = EmuLib(cfg)
    n = 70
    _, x0, pars = ensemble("lorenz63", n)
    sp = Solver(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, max_steps=333)
    dt, rng = np.full(n, sp.dt), seed_states(1, n)
    want = lib.features((0.0, 50.0), x0, pars, sp, Observer(), dt, rng)
    got = lib.features_chunked((0.0, 50.0), x0, pars, sp, Observer(), dt, rng, budgets=(100, 100, 100, 100, 100))
    got.pop("launches")
    assert_bit_equal(got, want, "max_steps cut-off")
    assert got["steps"].max() == 333


@pytest.mark.parametrize("observer", ["basicall", "localmax", "nhood2", "thresh2"])
@pytest.mark.parametrize("batch", [False, True])
def test_extents_in_shared_memory_placement_matches_oracle(observer, batch):
    """observers.cuh, CLODE_EXT_SMEM: the extents / means of the multi-variable observers (and thresh2's thresholds) in a
    shared-memory array instead of registers — what the runtime selects for features kernels that spill (C3) — and its
    opt-in variant that loads all words before it compares (CLODE_EXT_BATCH): both bit-identical to the oracle.  The
    emulation runs the threads of a block one after the other, so `__shared__` is a per-host-thread static array."""
    ns = 2 if observer in ("localmax", "nhood2", "thresh2") else 0
    n = 7
    ts, x0, pars = ensemble("lactotroph", n)
    ts = (ts[0], ts[1] / 5)
    sp = Solver(dt=0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=100000, max_store=90, nout=7)
    op = Observer(max_event_count=30, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    cfg = Config("lactotroph", "bs23", observer, ns, math="pm")
    defs = ("-DCLODE_EXT_SMEM", "-DCLODE_EMU_EXT_SMEM", "-DCLODE_BLOCK=128") + (("-DCLODE_EXT_BATCH",) if batch else ())
    A, B = EmuLib(cfg, extra_defs=defs), restate.OracleLib(cfg)
    got, want = run_oracle(A, "features", ts, x0, pars, sp, op), run_oracle(B, "features", ts, x0, pars, sp, op)
    assert_bit_equal(got, want, cfg.tag + (" batched" if batch else " in shared memory"))
    # and a continued call (the record travels through the persistent layout in between)
    r2a = A.features((ts[1], 2 * ts[1]), got["xf"], pars, sp, op, got["dt"], got["rng"], initialize=False)
    r2b = B.features((ts[1], 2 * ts[1]), want["xf"], pars, sp, op, want["dt"], want["rng"], initialize=False)
    assert_bit_equal(r2a, r2b, cfg.tag + " continued")


def test_extents_in_shared_memory_single_precision_and_stochastic():
    """the same placement (with the batched loads) in single precision, and for the stochastic stepper with all-variable
    extents: bit-identical to the oracle"""
    defs = ("-DCLODE_EXT_SMEM", "-DCLODE_EMU_EXT_SMEM", "-DCLODE_BLOCK=128", "-DCLODE_EXT_BATCH")
    cfg = Config("vanderpol", "dopri5", "thresh2", single=True)
    ts, x0, pars = ensemble("vanderpol", 6)
    sp = Solver(dt=0.1, dtmax=1.0, reltol=1e-3, max_steps=20000)
    op = Observer(x_up_threshold=0.3, x_down_threshold=0.2)
    A, B = EmuLib(cfg, extra_defs=defs), restate.OracleLib(cfg)
    assert_bit_equal(run_oracle(A, "features", ts, x0, pars, sp, op), run_oracle(B, "features", ts, x0, pars, sp, op), "single precision")
    cfg = Config("lactotroph_noise", "seuler", "basicall", math="pm")
    ts, x0, pars = ensemble("lactotroph_noise", 5)
    ts = (ts[0], ts[1] / 10)
    sp = Solver(dt=0.05, dtmax=1.0, max_steps=100000)
    A, B = EmuLib(cfg, extra_defs=defs), restate.OracleLib(cfg)
    assert_bit_equal(run_oracle(A, "features", ts, x0, pars, sp, Observer(), seed=5), run_oracle(B, "features", ts, x0, pars, sp, Observer(), seed=5),
                     "stochastic Euler, basicall")
