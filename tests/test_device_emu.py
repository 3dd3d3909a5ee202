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
from oracle.common import Config, Observer, Solver
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
