"""Pin the C restatement oracle (oracle/restate) against oracle/_ref — the reference's own
kernel sources compiled as host C — bit for bit, and pin both against the committed golden
vectors.  Also checks that clode_b200/models/*.cl are arithmetic-identical to the reference's
own RHS files.  CPU only."""
import os

import numpy as np
import pytest

from golden_cases import CASES, case_inputs
from oracle import ref, restate
from oracle.common import REFERENCE_ROOT, REFERENCE_RHS, MODELS, Config, Observer, Solver, seed_states
from problems import ensemble
from util import assert_bit_equal, run_oracle


def _have_ref(cfg):
    return os.path.exists(ref.so_path(cfg)) or ref.reference_available()


MATRIX = []
for model, steppers in [("lorenz63", ["euler", "heun", "rk4", "bs23", "dopri5"]), ("lactotroph", ["bs23", "dopri5"]),
                        ("vanderpol", ["rk4", "dopri5"])]:
    for st in steppers:
        for ob in ["basic", "basicall", "localmax", "nhood1", "nhood2", "thresh2"]:
            if model != "lorenz63" and ob in ("basic", "basicall"):
                continue
            MATRIX.append((model, st, ob))


@pytest.mark.parametrize("model,stepper,observer", MATRIX)
@pytest.mark.parametrize("math", ["libm", "pm"])
def test_restatement_matches_reference_kernels(model, stepper, observer, math):
    if math == "pm" and stepper in ("euler", "heun"):
        pytest.skip("fixed-step methods without transcendental calls: covered by the libm flavour")
    ns = 2 if observer in ("localmax", "nhood2", "thresh2") else 0
    cfg = Config(model, stepper, observer, ns, math=math)
    if not _have_ref(cfg):
        pytest.skip("oracle/_ref not built and reference tree absent")
    A, B = ref.RefLib(cfg), restate.OracleLib(cfg)
    n = 6
    ts, x0, pars = ensemble(model, n)
    ts = (ts[0], ts[1] / 5)
    fixed = stepper in ("euler", "heun", "rk4")
    sp = Solver(dt=0.01 if fixed else 0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=100000, max_store=150, nout=7)
    op = Observer(max_event_count=50, max_event_timestamps=ns, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    kinds = ["features"] + (["transient", "trajectory"] if observer == "basic" else [])
    for kind in kinds:
        assert_bit_equal(run_oracle(B, kind, ts, x0, pars, sp, op, seed=7), run_oracle(A, kind, ts, x0, pars, sp, op, seed=7),
                         f"{cfg.tag} {kind}")
    # continuation of the observer state across a second call (features(false) path)
    ra, rb = run_oracle(A, "features", ts, x0, pars, sp, op), run_oracle(B, "features", ts, x0, pars, sp, op)
    ts2 = (ts[1], ts[1] + (ts[1] - ts[0]))
    ra2 = A.features(ts2, ra["xf"], pars, sp, op, ra["dt"], ra["rng"], initialize=False)
    rb2 = B.features(ts2, rb["xf"], pars, sp, op, rb["dt"], rb["rng"], initialize=False)
    assert_bit_equal(rb2, ra2, f"{cfg.tag} continued features")


def test_stochastic_stream_matches_reference_kernels():
    cfg = Config("lactotroph_noise", "seuler", "basicall")
    if not _have_ref(cfg):
        pytest.skip("oracle/_ref not built and reference tree absent")
    A, B = ref.RefLib(cfg), restate.OracleLib(cfg)
    n = 16
    ts, x0, pars = ensemble("lactotroph_noise", n)
    sp = Solver(dt=0.01, max_steps=100000)
    ra, rb = run_oracle(A, "features", (0.0, 20.0), x0, pars, sp, Observer(), seed=1), \
        run_oracle(B, "features", (0.0, 20.0), x0, pars, sp, Observer(), seed=1)
    assert_bit_equal(rb, ra, "seuler features")
    assert not np.array_equal(ra["rng"], seed_states(1, n)), "the stream must have advanced"


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_matches_golden(name, golden):
    """committed vectors generated from the reference's own kernels (tests/golden/make_golden.py)"""
    case = CASES[name]
    cfg = Config(case["model"], case["stepper"], case.get("observer", "basic"), case.get("n_store", 0), math="pm")
    lib = restate.OracleLib(cfg)
    ts, x0, pars, sp, op, n = case_inputs(case)
    r = run_oracle(lib, case["kind"], ts, x0, pars, sp, op, seed=case.get("seed", 1))
    want = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/") and "/cont_" not in k}
    assert_bit_equal(r, want, name)
    if case.get("continue"):
        ts2 = (ts[1], ts[1] + (ts[1] - ts[0]))
        r2 = lib.features(ts2, r["xf"], pars, sp, op, r["dt"], r["rng"], initialize=False)
        want2 = {k.split("/cont_", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/cont_")}
        assert_bit_equal(r2, want2, name + " (continued)")


@pytest.mark.parametrize("model", sorted(REFERENCE_RHS))
def test_model_files_match_reference_rhs(model):
    """our RHS files must be arithmetic-identical to the reference's fixtures for the same system"""
    path = os.path.join(REFERENCE_ROOT, REFERENCE_RHS[model])
    if not (ref.reference_available() and os.path.exists(path)):
        pytest.skip("reference tree absent")
    stepper = "seuler" if model == "lactotroph_noise" else "rk4"
    ours = ref.RefLib(Config(model, stepper, "basicall"))
    theirs = ref.RefLib(Config(model + "_refrhs", stepper, "basicall", rhs_path=path, dims=MODELS[model]))
    n = 8
    ts, x0, pars = ensemble(model, n)
    sp = Solver(dt=0.01, max_steps=2000)
    assert_bit_equal(run_oracle(ours, "features", ts, x0, pars, sp, Observer()),
                     run_oracle(theirs, "features", ts, x0, pars, sp, Observer()), model)
