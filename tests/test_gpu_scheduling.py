"""Cost-sorted chunked execution of the adaptive time loops on the GPU (kernels.cuh "Scheduling", clode_rt.cpp run_loop):
the schedule changes WHEN and on WHICH lane an instance is integrated, never its arithmetic, so every output must be
bit-identical with the schedule switched off — in both tiers, on sorted and on shuffled parameter sets — and
bit-identical to the oracle in the bit-exact tier."""
import os

import numpy as np
import pytest

from oracle import restate
from oracle.common import MODELS, Config, Observer, Solver, seed_states
from problems import ensemble, rhs_source
from util import GpuRun, assert_bit_equal, run_oracle

pytestmark = pytest.mark.gpu


class _Sched:
    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self.old = os.environ.get("CLODE_SCHED")
        if self.value is None:
            os.environ.pop("CLODE_SCHED", None)
        else:
            os.environ["CLODE_SCHED"] = self.value

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("CLODE_SCHED", None)
        else:
            os.environ["CLODE_SCHED"] = self.old


CASES = [
    ("lorenz63", "dopri5", "basic", 0, (0.0, 40.0), 1.0),
    ("lorenz63", "dopri5", "localmax", 2, (0.0, 40.0), 1.0),
    ("lorenz63", "bs23", "nhood2", 2, (0.0, 20.0), 1.0),
    ("lactotroph", "bs23", "thresh2", 2, (0.0, 3000.0), 100.0),
    ("vanderpol", "dopri5", "nhood1", 0, (0.0, 60.0), 1.0),
]


@pytest.mark.parametrize("model, stepper, observer, n_store, tspan, dtmax", CASES)
@pytest.mark.parametrize("bit_exact", [True, False])
@pytest.mark.parametrize("shuffle", [False, True])
def test_cuda_schedule_does_not_change_results(rt, model, stepper, observer, n_store, tspan, dtmax, bit_exact, shuffle):
    n = 9 * 128 + 57
    _, x0, pars = ensemble(model, n)
    nv, npar = MODELS[model][:2]
    if shuffle:
        perm = np.random.default_rng(3).permutation(n)
        x0 = x0.reshape(nv, n)[:, perm].ravel()
        pars = pars.reshape(npar, n)[:, perm].ravel()
    sp = Solver(dt=0.01, dtmax=dtmax, abstol=1e-6, reltol=1e-5, max_steps=1000000)
    op = Observer(max_event_count=200, max_event_timestamps=n_store, x_up_threshold=0.3, x_down_threshold=0.2)
    g = GpuRun(rt, model, stepper, observer, n_store=n_store, bit_exact=bit_exact)
    results = {}
    for name, sched in (("off", "off"), ("default", None), ("tiny", "3,5,0.3,8"), ("one_round", "64,1,0.5,64")):
        with _Sched(sched):
            g.setup(tspan, x0, pars, sp, op, seed=1)
            results[name] = g.features()
            launches = g.sim.launch_count()
            g.setup(tspan, x0, pars, sp, op, seed=1)
            results[name + "_transient"] = g.transient()
    assert launches > 4  # the schedule really ran as several launches
    for name in ("default", "tiny", "one_round"):
        assert_bit_equal(results[name], results["off"], f"features, schedule {name} vs off ({model}/{stepper}/{observer})")
        assert_bit_equal(results[name + "_transient"], results["off_transient"], f"transient, schedule {name} vs off")
    if bit_exact and not shuffle:
        lib = restate.OracleLib(Config(model, stepper, observer, n_store, math="pm"))
        o = run_oracle(lib, "features", tspan, x0, pars, sp, op, seed=1)
        assert_bit_equal({k: results["default"][k] for k in ("F", "xf", "tf", "dt")}, o, "scheduled features vs oracle")
    g.close()


def test_cuda_schedule_continuation_and_two_pass_cost_reuse(rt):
    """features(initialize=1) of a two-pass observer sorts the features pass by the warm-up's exact step counts; a continued
    call (initialize=0) goes through pilot + rounds again; both equal the unscheduled run bit for bit"""
    n = 5 * 128 + 3
    _, x0, pars = ensemble("lactotroph", n)
    sp = Solver(dt=0.1, dtmax=100.0, abstol=1e-6, reltol=1e-4, max_steps=1000000)
    op = Observer(max_event_count=1000, x_up_threshold=0.3, x_down_threshold=0.2)
    out = {}
    for name, sched in (("off", "off"), ("on", None)):
        with _Sched(sched):
            g = GpuRun(rt, "lactotroph", "bs23", "thresh2", bit_exact=False)
            g.setup((0.0, 2000.0), x0, pars, sp, op, seed=1)
            first = g.features(1)
            g.sim.shift_x0()
            g.sim.set_tspan(2000.0, 4000.0)
            second = g.features(0)
            out[name] = (first, second)
            g.close()
    assert_bit_equal(out["on"][0], out["off"][0], "two-pass features, scheduled vs off")
    assert_bit_equal(out["on"][1], out["off"][1], "continued features, scheduled vs off")


@pytest.mark.parametrize("model, stepper, observer, tspan, dtmax", [("lactotroph", "bs23", "thresh2", (0.0, 1500.0), 100.0),
                                                                     ("lorenz63", "dopri5", "basic", (0.0, 20.0), 1.0),
                                                                     ("lactotroph_noise", "seuler", "basicall", (0.0, 20.0), 1.0)])
def test_cuda_literals_through_the_constant_bank_do_not_change_results(rt, model, stepper, observer, tspan, dtmax):
    """csrc/rt/ptx_pass.hpp hoist_f64_immediates: double literals become constant-bank operands instead of UMOV pairs —
    same values, same operations, so the production build's results are bit-identical with the pass off"""
    n = 3 * 128 + 11
    _, x0, pars = ensemble(model, n)
    sp = Solver(dt=0.01, dtmax=dtmax, abstol=1e-6, reltol=1e-5, max_steps=1000000)
    op = Observer(max_event_count=200, x_up_threshold=0.3, x_down_threshold=0.2)
    out = {}
    for name, env in (("hoist", "1"), ("plain", "0")):
        old = os.environ.get("CLODE_IMM_HOIST")
        os.environ["CLODE_IMM_HOIST"] = env
        try:
            g = GpuRun(rt, model, stepper, observer, bit_exact=False)
            log = g.sim.build_log() if hasattr(g.sim, "build_log") else ""
            g.setup(tspan, x0, pars, sp, op, seed=1)
            out[name] = g.features()
            g.close()
        finally:
            if old is None:
                os.environ.pop("CLODE_IMM_HOIST", None)
            else:
                os.environ["CLODE_IMM_HOIST"] = old
    assert_bit_equal(out["hoist"], out["plain"], f"{model}: literals through the constant bank")
