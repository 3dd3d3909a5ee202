"""helpers shared by the CPU and GPU test modules"""
from __future__ import annotations

import numpy as np

from oracle.common import MODELS, Config, Observer, Solver, seed_states
from problems import ensemble, rhs_source


def assert_bit_equal(got: dict, want: dict, what: str, keys=None):
    """bit-for-bit equality of every array (NaNs in the same places count as equal)"""
    for k in (keys or want.keys()):
        if k == "rows" or k not in got:
            continue
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, f"{what}: {k} shape {a.shape} != {b.shape}"
        if a.dtype.kind == "f":
            same = (a.view(np.uint64 if a.dtype == np.float64 else np.uint32) ==
                    b.view(np.uint64 if b.dtype == np.float64 else np.uint32)) | (np.isnan(a) & np.isnan(b)) | (a == b)
        else:
            same = a == b
        if not same.all():
            bad = np.flatnonzero(~same)
            raise AssertionError(f"{what}: {k} differs at {bad.size}/{a.size} places, first {bad[:4]}: "
                                 f"{a[bad[:4]]} vs {b[bad[:4]]}")


def run_oracle(lib, kind, ts, x0, pars, sp, op, seed=1, n=None):
    n = n or len(x0) // lib.n_var
    dt, rng = np.full(n, sp.dt), seed_states(seed, n)
    if kind == "transient":
        return lib.transient(ts, x0, pars, sp, dt, rng)
    if kind == "features":
        return lib.features(ts, x0, pars, sp, op, dt, rng)
    return lib.trajectory(ts, x0, pars, sp, dt, rng)


class GpuRun:
    """Drive the CUDA kernels through the C ABI (clode_b200._rt.Sim) with oracle-style arguments."""

    def __init__(self, rt, model, stepper, observer="basic", n_store=0, bit_exact=True, single=False,
                 f_var=0, e_var=0, work_queue=False, block=0, staged=False, obs_smem=False):
        nv, npar, na, nw = MODELS[model]
        self.rt = rt
        self.prog = rt.Program(rhs_source(model), stepper, nv, npar, na, nw, observer=observer, bit_exact=bit_exact,
                               n_store_events=n_store, f_var_ix=f_var, e_var_ix=e_var, single_precision=single,
                               work_queue=work_queue, block_size=block, staged_trajectory=staged, observer_in_shared=obs_smem)
        self.sim = rt.Sim(self.prog)
        self.n_store = n_store

    def setup(self, ts, x0, pars, sp: Solver, op: Observer | None, seed=1, rng=None, dt=None):
        s = self.sim
        s.set_solver_params(sp.dt, sp.dtmax, sp.abstol, sp.reltol, sp.max_steps, sp.max_store, sp.nout)
        if op is not None:
            s.set_observer_params(op.e_var_ix, op.f_var_ix, op.max_event_count, self.n_store, op.min_amp, op.min_imi,
                                  op.nhood_radius, op.x_up_threshold, op.x_down_threshold, op.dx_up_threshold,
                                  op.dx_down_threshold, op.eps_dx)
        s.set_tspan(*ts)
        s.set_problem(x0, pars)
        if rng is not None:
            s.set_rng_state(rng)
        else:
            s.seed_rng(seed)
        if dt is not None:
            s.set_dt(dt)

    def common(self):
        s = self.sim
        return dict(xf=s.get_xf(), tf=s.get_tf(), dt=s.get_dt(), rng=s.get_rng_state(), steps=s.get_steps())

    def transient(self):
        self.sim.transient()
        return self.common()

    def features(self, initialize=1):
        self.sim.features(initialize)
        return dict(self.common(), F=self.sim.get_f())

    def trajectory(self):
        self.sim.trajectory()
        return dict(self.common(), **self.sim.get_trajectory())

    def run(self, kind):
        return getattr(self, kind)()

    def close(self):
        self.sim.close()


def cast_like(result: dict, real):
    """the C ABI returns doubles; oracle results in single precision are float32"""
    return {k: (np.asarray(v, dtype=np.float64) if np.asarray(v).dtype == np.float32 else v) for k, v in result.items()}
