"""Known-answer tests the REFERENCE's own test-suite holds for the hot path, applied to the CPU
oracles (restatement and, when built, oracle/_ref).  Each test cites the reference test it mirrors.
CPU only."""
import os
from math import exp, log, pi, sqrt

import numpy as np
import pytest

from oracle import ref, restate
from oracle.common import Config, Observer, Solver, seed_states


def _libs(cfg):
    libs = [("restate", restate.OracleLib(cfg))]
    if os.path.exists(ref.so_path(cfg)) or ref.reference_available():
        libs.append(("_ref", ref.RefLib(cfg)))
    return libs


@pytest.mark.parametrize("single", [False, True])
def test_ornl_thompson_a1_exact_solution(single):
    """test/test_ornl_thompson_a1.py:16-46 — rk4, dt=1e-3, every stored point to 5 significant digits"""
    cfg = Config("thompson_a1", "rk4", single=single)
    H = 10.0
    x0, pars = np.zeros(2), np.array([0.25, 8.0, 2.0, H])
    sp = Solver(dt=0.001, dtmax=0.001, max_store=20000, max_steps=20000)
    for name, lib in _libs(cfg):
        r = lib.trajectory((0.0, H / 2.0), x0, pars, sp, np.full(1, sp.dt), seed_states(1, 1))
        ns = int(r["n_stored"][0]) + 1
        assert ns in (5001, 5002)  # `<=` loop: t accumulates to just above/below 5.0 after 5000 steps
        t = r["t"][:ns].astype(np.float64)
        x = r["x"].reshape(r["rows"], 2)[:ns].astype(np.float64)
        y1 = 4.0 * (t + np.exp(-8.0 * t) / 8.0 - 1.0 / 8.0)
        y2 = 4.0 * (1.0 - np.exp(-8.0 * t))
        for got, want in ((x[1:, 0], y1[1:]), (x[1:, 1], y2[1:])):
            # np.testing.assert_approx_equal(significant=5): |got-want| < 10^-(5-1) * 10^floor(log10(scale))
            scale = 10.0 ** np.floor(np.log10(0.5 * (np.abs(got) + np.abs(want))))
            assert np.all(np.abs(got - want) / scale < 10.0 ** -4), name


def test_thresh2_event_times_on_sine():
    """test/test_features.py:24-76 — event count 2; up at pi/4, 9pi/4; down at 3pi/2, 7pi/2 (atol 0.01)"""
    cfg = Config("sine_drive", "rk4", "thresh2", n_store_events=3)
    sp = Solver(dt=0.001, dtmax=0.001, max_steps=10000000)
    op = Observer(e_var_ix=0, f_var_ix=0, max_event_count=100, max_event_timestamps=3, min_amp=0.5,
                  x_up_threshold=(2 + sqrt(2)) / 4, x_down_threshold=0.001, dx_down_threshold=0.001,
                  dx_up_threshold=0.001)
    for name, lib in _libs(cfg):
        r = lib.features((0.0, 4 * pi), np.zeros(1), np.ones(1), sp, op, np.full(1, sp.dt), seed_states(1, 1))
        F = r["F"]
        base = 18 + 5 * 1 + 3 * 3
        up, down = F[base:base + 6:2], F[base + 1:base + 6:2]
        assert int(F[base + 6]) == 2, name  # "event count"
        assert np.isclose(up[0], pi / 4, atol=0.01) and np.isclose(up[1], 9 * pi / 4, atol=0.01), (name, up)
        assert np.isclose(down[0], 3 * pi / 2, atol=0.01) and np.isclose(down[1], 7 * pi / 2, atol=0.01), (name, down)
        assert up[2] == 0.0 and down[2] == 0.0


@pytest.mark.parametrize("observer", ["basicall", "localmax", "nhood1", "nhood2", "thresh2"])
def test_aux_extents_and_means(observer):
    """test/test_aux_values.py:27-91 — aux max/min/mean plumbing (constant aux must be exact).
    The reference test parametrises over observers but always builds thresh2; here each observer runs."""
    cfg = Config("sine_drive", "rk4", observer)
    sp = Solver(dt=0.1, dtmax=0.1, max_steps=10000000)
    op = Observer(x_up_threshold=0.3, x_down_threshold=0.2)
    lead = {"basicall": 0, "localmax": 6, "nhood1": 6, "nhood2": 6, "thresh2": 18}[observer]
    per_var = 7 if observer == "nhood2" else 5
    for name, lib in _libs(cfg):
        r = lib.features((0.0, 400 * pi), np.zeros(1), np.ones(1), sp, op, np.full(1, sp.dt), seed_states(1, 1))
        aux = r["F"][lead + per_var:lead + per_var + 9].reshape(3, 3)  # rows: xp1, pos, neg; cols: max, min, mean
        assert aux[0, 2] == pytest.approx(1.0, abs=1e-2) and aux[0, 1] == pytest.approx(0.0, abs=1e-2)
        assert aux[0, 0] == pytest.approx(2.0, abs=1e-2)
        assert tuple(aux[1]) == (1.0, 1.0, 1.0), name
        assert aux[2, 0] == -2.0 and aux[2, 1] == -2.0, name


def _vdp_period(mu):
    """test/test_vdp.py:16-30"""
    if mu < 0:
        return 0.0
    if mu < 2:
        return 2 * pi * (1 + mu ** 2 / 16)
    return min(2 * pi * (1 + mu ** 2 / 16), (3 - 2 * log(2)) * mu + 3 * 2.2338 / mu ** (1 / 3.0))


@pytest.mark.parametrize("single", [True, False])
def test_van_der_pol_period(single):
    """test/test_vdp.py:50-91 — dopri5 + thresh2, `max period` vs the analytic approximation, rtol 1 %, atol 1"""
    cfg = Config("vanderpol", "dopri5", "thresh2", single=single)
    mus = np.array([-1, 0, 0.01, 0.1, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 5, 6], dtype=float)
    n = len(mus)
    sp = Solver(dt=0.1, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=20000, max_store=20000)
    op = Observer(max_event_count=100, x_up_threshold=0.3, x_down_threshold=0.2)
    for name, lib in _libs(cfg):
        r = lib.features((0.0, 1000.0), np.ones(2 * n), mus, sp, op, np.full(n, sp.dt), seed_states(1, n))
        period = r["F"][:n].astype(np.float64)  # feature 0 = "max period"
        for k, mu in enumerate(mus):
            assert np.isclose(period[k], _vdp_period(mu), rtol=0.01, atol=1), (name, mu, period[k])


def test_lactotroph_sample_final_state():
    """samples/test_outputs_cpp.md:10 — lactotroph, rk4 dt=0.1, t in [0,1000], x0 = 0, p = (1.5, 3, 1):
    pasted output `xf = -64.7058 0.022137 6.25581e-06 0.209086` (single precision, 'some variability')."""
    cfg = Config("lactotroph", "rk4", single=True)
    sp = Solver(dt=0.1, dtmax=1.0, max_steps=10000000)
    for name, lib in _libs(cfg):
        r = lib.transient((0.0, 1000.0), np.zeros(4), np.array([1.5, 3.0, 1.0]), sp, np.full(1, sp.dt), seed_states(1, 1))
        assert np.allclose(r["xf"], [-64.7058, 0.022137, 6.25581e-06, 0.209086], rtol=2e-3), (name, r["xf"])
