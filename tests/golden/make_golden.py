"""Generate tests/golden/golden_ref.npz from oracle/_ref — the REFERENCE's own kernel sources
compiled as host C (oracle/ref.py) — for a fixed list of small, seeded cases.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The .npz is committed; tests compare the C restatement oracle, the host-emulated device code
and the CUDA kernels against it.  Cases use the portable-math flavour ("pm", bit-exact tier)
so the expected bits do not depend on the libm of the machine that runs the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref  # noqa: E402
from oracle.common import Config, seed_states  # noqa: E402
from golden_cases import CASES, case_inputs  # noqa: E402


def main():
    out = {}
    for name, case in CASES.items():
        cfg = Config(case["model"], case["stepper"], case.get("observer", "basic"),
                     case.get("n_store", 0), math="pm")
        lib = ref.RefLib(cfg)
        ts, x0, pars, sp, op, n = case_inputs(case)
        dt, rng = np.full(n, sp.dt), seed_states(case.get("seed", 1), n)
        kind = case["kind"]
        if kind == "transient":
            r = lib.transient(ts, x0, pars, sp, dt, rng)
        elif kind == "features":
            r = lib.features(ts, x0, pars, sp, op, dt, rng)
            if case.get("continue"):
                ts2 = (ts[1], ts[1] + (ts[1] - ts[0]))
                r2 = lib.features(ts2, r["xf"], pars, sp, op, r["dt"], r["rng"], initialize=False)
                for k, v in r2.items():
                    out[f"{name}/cont_{k}"] = v
        else:
            r = lib.trajectory(ts, x0, pars, sp, dt, rng)
            r.pop("rows")
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        print(name, {k: v.shape for k, v in r.items()})
    path = os.path.join(HERE, "golden_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
