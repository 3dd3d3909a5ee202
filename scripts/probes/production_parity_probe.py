#!/usr/bin/env python
"""How far is the PRODUCTION tier (FMA contraction, fast controller root, Newton reciprocals, integral means) from
the reference arithmetic (oracle, libm flavour, no contraction) at the BASELINE sizes?  Prints, per workload and per
parameter band, the fraction of sampled instances with IDENTICAL accepted-step counts / event counts and the largest
relative feature deviation.  Numbers from this script set the bounds asserted in tests/test_gpu_production_parity.py.

  python scripts/probes/production_parity_probe.py [C2 C3 C5 ...] [--k 2048]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["C2", "C3"])
    ap.add_argument("--k", type=int, default=2048)
    ap.add_argument("--npts", type=int, default=0)
    args = ap.parse_args()

    import bench
    from clode_b200 import _rt, build, sharding
    from clode_b200.models import MODELS, rhs_source
    from oracle import restate
    from oracle.common import Config, Observer, Solver

    build.build_runtime()
    if "C5" in args.workloads:
        args.workloads.remove("C5")
        n = args.npts or (1 << 18)
        w = bench.workload("C5", n, np.arange(n))
        nv, npar, na, nw = MODELS[w["model"]]
        sub = np.sort(np.random.default_rng(7).choice(n, 256, replace=False))
        x0s, ps = sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub)
        prog = _rt.Program(rhs_source(w["model"]), w["stepper"], nv, npar, na, nw, kernels=_rt.KERNEL_TRAJECTORY)
        sim = _rt.Sim(prog)
        sim.set_solver_params(**w["solver"])
        sim.set_tspan(*w["tspan"])
        sim.set_problem(x0s, ps)
        sim.seed_rng(1)
        sim.trajectory()
        tr = sim.get_trajectory()
        rows = w["solver"]["max_store"]
        lib = restate.OracleLib(Config(w["model"], w["stepper"], math="libm"))
        sp = Solver(**w["solver"])
        o = lib.trajectory(w["tspan"], x0s, ps, sp, np.full(sub.size, sp.dt), sharding.seed_states_for(1, n, sub))
        m = sub.size
        xg = np.asarray(tr["x"]).reshape(-1, nv, m)[:rows]
        xo = np.asarray(o["x"]).reshape(-1, nv, m)[:rows]
        scale = np.abs(xo).max(axis=(0, 2), keepdims=True)
        dev = np.abs(xg - xo) / scale
        print(json.dumps({"workload": "C5", "sampled": m, "n_stored_identical": bool(np.array_equal(tr["n_stored"], o["n_stored"])),
                          "t_identical": bool(np.array_equal(np.asarray(tr["t"])[:rows * m], np.asarray(o["t"])[:rows * m])),
                          "max_dev_rows_0_200": float(dev[:200].max()), "max_dev_rows_0_1000": float(dev[:1000].max()),
                          "max_dev_all": float(dev.max()), "median_final_dev": float(np.median(dev[-1])),
                          "p99_all": float(np.quantile(dev, 0.99))}), flush=True)
        sim.close()
    for name in args.workloads:
        n = args.npts or {"C4": 1 << 22, "C5": 1 << 18}.get(name, 1 << 20)
        w = bench.workload(name, n, np.arange(n))
        nv, npar, na, nw = MODELS[w["model"]]
        prog = _rt.Program(rhs_source(w["model"]), w["stepper"], nv, npar, na, nw, observer=w["observer"], kernels=_rt.KERNEL_FEATURES)
        sim = _rt.Sim(prog)
        sim.set_solver_params(**w["solver"])
        sim.set_observer_params(**w["observer_params"])
        sim.set_tspan(*w["tspan"])
        sim.set_problem(w["x0"], w["pars"])
        sim.seed_rng(1)
        sim.features(1)
        nf = sim.n_features()
        F = sim.get_f().reshape(nf, n)
        steps = sim.get_steps().astype(np.int64)
        sub = np.sort(np.random.default_rng(7).choice(n, args.k, replace=False))
        lib = restate.OracleLib(Config(w["model"], w["stepper"], w["observer"], math="libm"))
        sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
        o = lib.features(w["tspan"], sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub), sp, op,
                         np.full(sub.size, sp.dt), sharding.seed_states_for(1, n, sub), nthreads=os.cpu_count())
        Fo = o["F"].reshape(nf, sub.size)
        Fg = F[:, sub]
        step_row = {"basic": 5}.get(w["observer"], nf - (1 if w["observer"] in ("basicall", "localmax") else 4))
        so, sg = Fo[step_row].astype(np.int64), Fg[step_row].astype(np.int64)
        same = so == sg
        scale = np.maximum(np.abs(Fo).max(axis=1, keepdims=True), 1e-300)
        dev = np.abs(Fg - Fo) / scale  # deviation relative to the feature's range over the sample
        out = {"workload": name, "sampled": int(sub.size), "identical_step_count_fraction": float(same.mean()),
               "max_abs_step_diff": int(np.abs(so - sg).max()), "rel_step_diff_max": float((np.abs(so - sg) / np.maximum(so, 1)).max()),
               "total_steps_gpu": int(steps.sum())}
        if name.startswith("C2"):
            r = w["pars"].reshape(npar, n)[0, sub]
            for lo, hi, label in ((0.0, 1.0, "r<1 (origin stable)"), (1.0, 13.9, "1<r<13.9 (fixed points, no transient chaos)"),
                                  (13.9, 24.06, "13.9<r<24.06 (transient chaos)"), (24.06, 1e9, "r>24.06 (chaotic)")):
                m = (r >= lo) & (r < hi)
                if m.any():
                    out[label] = {"instances": int(m.sum()), "identical_steps": float(same[m].mean()),
                                  "max_step_diff": int(np.abs(so - sg)[m].max()),
                                  "max_feature_dev_rel_range": float(dev[:step_row][:, m].max())}
        else:
            ev_row = step_row - 1
            if w["observer"] in ("thresh2", "nhood2", "localmax"):
                eo, eg = Fo[ev_row].astype(np.int64), Fg[ev_row].astype(np.int64)
                out["identical_event_count_fraction"] = float((eo == eg).mean())
                out["max_event_diff"] = int(np.abs(eo - eg).max())
                both = (eo == eg) & same
                out["identical_steps_and_events_fraction"] = float(both.mean())
                if both.any():
                    out["max_feature_dev_rel_range_where_counts_identical"] = float(np.nanmax(dev[:, both]))
            out["max_feature_dev_rel_range"] = float(np.nanmax(dev))
            out["per_feature_dev"] = [float(x) for x in np.nanmax(dev, axis=1)]
        print(json.dumps(out), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
