"""The fixed, seeded cases behind tests/golden/golden_ref.npz (see tests/golden/make_golden.py)."""
from __future__ import annotations

import numpy as np

from problems import ensemble
from oracle.common import Observer, Solver

_ADAPT = dict(dt=0.1, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=100000, max_store=120, nout=3)
_FIXED = dict(dt=0.01, dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=100000, max_store=120, nout=25)
_OBS = dict(max_event_count=40, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)

CASES = {
    # steppers (transient kernel), Lorenz and Van der Pol
    "tr_euler": dict(model="lorenz63", stepper="euler", kind="transient", solver=_FIXED, t_scale=0.1),
    "tr_heun": dict(model="lorenz63", stepper="heun", kind="transient", solver=_FIXED, t_scale=0.1),
    "tr_rk4_vdp": dict(model="vanderpol", stepper="rk4", kind="transient", solver=_FIXED, t_scale=0.2),
    "tr_bs23": dict(model="lorenz63", stepper="bs23", kind="transient", solver=_ADAPT, t_scale=0.1),
    "tr_dopri5": dict(model="lorenz63", stepper="dopri5", kind="transient", solver=_ADAPT, t_scale=0.1),
    # observers (initializeObserver + features kernels)
    "ft_basic": dict(model="lorenz63", stepper="dopri5", observer="basic", kind="features", solver=_ADAPT, t_scale=0.2),
    "ft_basicall": dict(model="lactotroph", stepper="dopri5", observer="basicall", kind="features", solver=_ADAPT, t_scale=0.5),
    "ft_localmax": dict(model="lorenz63", stepper="dopri5", observer="localmax", n_store=3, kind="features",
                        solver=_ADAPT, t_scale=0.2, **{"continue": True}),
    "ft_nhood1": dict(model="lactotroph", stepper="bs23", observer="nhood1", kind="features", solver=_ADAPT, t_scale=1.0),
    "ft_nhood2": dict(model="lactotroph", stepper="dopri5", observer="nhood2", n_store=2, kind="features",
                      solver=_ADAPT, t_scale=1.0),
    "ft_thresh2": dict(model="lactotroph", stepper="bs23", observer="thresh2", n_store=4, kind="features",
                       solver=_ADAPT, t_scale=1.0, **{"continue": True}),
    "ft_thresh2_vdp": dict(model="vanderpol", stepper="dopri5", observer="thresh2", kind="features", solver=_ADAPT,
                           t_scale=0.5),
    "ft_terminal": dict(model="lactotroph", stepper="rk4", observer="thresh2", kind="features",
                        solver=dict(_FIXED, dt=0.05), t_scale=1.0, obs=dict(_OBS, max_event_count=3)),
    # stochastic Euler: per-instance RNG streams
    "ft_seuler": dict(model="lactotroph_noise", stepper="seuler", observer="basicall", kind="features",
                      solver=_FIXED, t_scale=0.2, seed=1),
    "tr_seuler": dict(model="lactotroph_noise", stepper="seuler", kind="transient", solver=_FIXED, t_scale=0.1, seed=5),
    # trajectory kernel, including the max_store cut-off (row index max_store is written)
    "tj_rk4": dict(model="chay_keizer", stepper="rk4", kind="trajectory", solver=dict(_FIXED, dt=0.5, nout=4), t_scale=0.2),
    "tj_dopri5": dict(model="chay_keizer", stepper="dopri5", kind="trajectory", solver=_ADAPT, t_scale=1.0),
    "tj_cut": dict(model="thompson_a1", stepper="rk4", kind="trajectory",
                   solver=dict(_FIXED, dt=0.001, nout=1, max_store=50), t_scale=1.0),
}

N_GOLDEN = 12


def case_inputs(case, n: int = N_GOLDEN):
    ts, x0, pars = ensemble(case["model"], n)
    ts = (ts[0], ts[0] + (ts[1] - ts[0]) * case.get("t_scale", 1.0))
    sp = Solver(**case["solver"])
    op = Observer(**case.get("obs", _OBS))
    op.max_event_timestamps = case.get("n_store", 0)
    return ts, x0, pars, sp, op, n
