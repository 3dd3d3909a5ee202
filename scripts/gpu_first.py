"""First GPU contact: parity of a few configurations against the C restatement oracle,
plus a rough Lorenz dopri5 timing.  Scratch diagnostics, not a test."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from problems import ensemble, rhs_source
from oracle.common import Config, Solver, Observer, seed_states, MODELS
from oracle import restate
from clode_b200 import _rt

print(_rt.device_info(0).name, _rt.device_info(0).multiprocessors)

def run_gpu(model, stepper, observer, n, sp, op, bit_exact, kind, ns=0, single=False, wq=False):
    nv, npar, na, nw = MODELS[model]
    prog = _rt.Program(rhs_source(model), stepper, nv, npar, na, nw, observer=observer, bit_exact=bit_exact,
                       n_store_events=ns, f_var_ix=op.f_var_ix, e_var_ix=op.e_var_ix, single_precision=single, work_queue=wq)
    sim = _rt.Sim(prog)
    ts, x0, pars = ensemble(model, n)
    sim.set_solver_params(sp.dt, sp.dtmax, sp.abstol, sp.reltol, sp.max_steps, sp.max_store, sp.nout)
    sim.set_observer_params(op.e_var_ix, op.f_var_ix, op.max_event_count, ns, op.min_amp, op.min_imi, op.nhood_radius,
                            op.x_up_threshold, op.x_down_threshold, op.dx_up_threshold, op.dx_down_threshold, op.eps_dx)
    sim.set_tspan(*ts)
    sim.set_problem(x0, pars)
    sim.seed_rng(7)
    if kind == "features":
        sim.features(1)
        out = dict(F=sim.get_f(), xf=sim.get_xf(), tf=sim.get_tf(), dt=sim.get_dt(), rng=sim.get_rng_state())
    elif kind == "transient":
        sim.transient()
        out = dict(xf=sim.get_xf(), tf=sim.get_tf(), dt=sim.get_dt(), rng=sim.get_rng_state())
    else:
        sim.trajectory()
        out = sim.get_trajectory(); out.update(xf=sim.get_xf(), tf=sim.get_tf(), dt=sim.get_dt())
    out["ms"] = sim.last_kernel_ms(); out["steps"] = sim.get_steps()
    out["info"] = sim.kernel_info({"features":2,"transient":1,"trajectory":4}[kind])
    sim.close()
    return out

def run_cpu(model, stepper, observer, n, sp, op, math, kind, ns=0):
    cfg = Config(model, stepper, observer, ns, math=math)
    L = restate.OracleLib(cfg)
    ts, x0, pars = ensemble(model, n)
    dt = np.full(n, sp.dt); rng = seed_states(7, n)
    op2 = Observer(**{**op.__dict__, "max_event_timestamps": ns})
    if kind == "features": return L.features(ts, x0, pars, sp, op2, dt, rng)
    if kind == "transient": return L.transient(ts, x0, pars, sp, dt, rng)
    return L.trajectory(ts, x0, pars, sp, dt, rng)

def compare(tag, g, c):
    bad = []
    for k in c:
        if k in ("rows",) or k not in g: continue
        a, b = np.asarray(g[k]), np.asarray(c[k])
        if a.shape != b.shape: bad.append((k, "shape", a.shape, b.shape)); continue
        if not np.array_equal(a, b, equal_nan=True):
            if a.dtype.kind == "f":
                rel = np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
                bad.append((k, int((a != b).sum()), float(rel)))
            else:
                bad.append((k, int((a != b).sum())))
    print(("EXACT " if not bad else "DIFF  ") + tag, bad[:6], flush=True)

n = 256
cases = [
 ("vanderpol","rk4","basic","transient"), ("lorenz63","dopri5","basic","features"), ("lorenz63","dopri5","localmax","features"),
 ("lorenz63","bs23","nhood2","features"), ("lorenz63","dopri5","nhood1","features"), ("lactotroph","bs23","thresh2","features"),
 ("lactotroph","dopri5","basicall","features"), ("lactotroph_noise","seuler","basicall","features"),
 ("chay_keizer","rk4","basic","trajectory"), ("chay_keizer","dopri5","basic","trajectory"), ("sine_drive","rk4","thresh2","features"),
 ("lorenz63","heun","basic","transient"), ("lorenz63","euler","basicall","features"),
]
for model, stepper, observer, kind in cases:
    fixed = stepper in ("euler","heun","rk4","seuler")
    sp = Solver(dt=(0.01 if fixed else 0.1), dtmax=10.0, abstol=1e-6, reltol=1e-4, max_steps=200000, max_store=300, nout=5)
    if model in ("lactotroph","chay_keizer") and fixed: sp.dt = 0.05
    op = Observer(max_event_count=50, x_up_threshold=0.3, x_down_threshold=0.2, nhood_radius=0.1)
    ns = 2 if observer in ("localmax","thresh2","nhood2") else 0
    g = run_gpu(model, stepper, observer, n, sp, op, True, kind, ns)
    c = run_cpu(model, stepper, observer, n, sp, op, "pm", kind, ns)
    compare(f"bitexact {model} {stepper} {observer} {kind} regs={g['info']['registers']} local={g['info']['local_bytes']} ms={g['ms']:.2f}", g, c)
    g = run_gpu(model, stepper, observer, n, sp, op, False, kind, ns)
    c = run_cpu(model, stepper, observer, n, sp, op, "libm", kind, ns)
    compare(f"fast     {model} {stepper} {observer} {kind} regs={g['info']['registers']} local={g['info']['local_bytes']} ms={g['ms']:.2f}", g, c)

# rough throughput: C2 Lorenz dopri5 basic, 2^20 instances
for wq in (False, True):
  for observer in ("basic", "localmax"):
    sp = Solver(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, max_steps=10000000)
    op = Observer(max_event_count=10000)
    for rep in range(2):
        g = run_gpu("lorenz63", "dopri5", observer, 1 << 20, sp, op, False, "features", wq=wq)
    tot = int(g["steps"].astype(np.int64).sum())
    print(f"C2 lorenz dopri5 {observer} wq={wq}: {g['ms']:.1f} ms, steps={tot:.3e}, {tot / g['ms'] * 1e3:.3e} steps/s, info={g['info']}", flush=True)
