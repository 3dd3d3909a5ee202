#!/usr/bin/env python
"""profiles/r02_sass_loop_histograms.txt: static SASS instruction mix of every benchmarked time loop (no GPU needed:
NVRTC + PTX pass + ptxas cross-compile for sm_100a).  usage: python scripts/make_sass_histograms.py [out file]"""
import os
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from clode_b200 import _rt, build  # noqa: E402
from clode_b200.models import MODELS, rhs_source  # noqa: E402


def main():
    build.build_runtime()
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02_sass_loop_histograms.txt")
    F, T, J = _rt.KERNEL_FEATURES, _rt.KERNEL_TRANSIENT, _rt.KERNEL_TRAJECTORY
    cases = [("C2 production", "lorenz63", "dopri5", "basic", F, "clode_features", 5, {}),
             ("C2 bit-exact tier", "lorenz63", "dopri5", "basic", F, "clode_features", 5, dict(bit_exact=True)),
             ("C2 localmax", "lorenz63", "dopri5", "localmax", F, "clode_features", 4, {}),
             ("C2t transient", "lorenz63", "dopri5", "basic", T, "clode_transient", 5, {}),
             ("C3 features", "lactotroph", "bs23", "thresh2", F, "clode_features", 4, {}),
             ("C3 warm-up", "lactotroph", "bs23", "thresh2", F, "clode_initialize_observer", 4, {}),
             ("C4", "lactotroph_noise", "seuler", "basicall", F, "clode_features", 4, {}),
             ("C5 rk4", "chay_keizer", "rk4", "basic", J, "clode_trajectory", 4, {})]
    with open(path, "w") as out:
        out.write("SASS instruction histograms of the time loops (scripts/sass_loop_hist.py on the cubins NVRTC + the PTX pass + ptxas\n"
                  "produce for sm_100a; the loop = the largest backward-branch span of the kernel; static counts, rare slow paths\n"
                  "included).  Regenerate: python scripts/make_sass_histograms.py\n\n")
        for tag, model, stepper, obs, kern, kname, m, kw in cases:
            nv, npar, na, nw = MODELS[model]
            prog = _rt.Program(rhs_source(model), stepper, nv, npar, na, nw, observer=obs, kernels=kern, min_blocks_per_sm=m, **kw)
            if obs == "thresh2":  # what clode_sim_build decides on the GPU for this program: observer extents in shared memory
                os.environ["CLODE_EXT_SMEM"] = "1"
            cubin, _ = _rt.compile_program(prog)
            os.environ.pop("CLODE_EXT_SMEM", None)
            with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
                f.write(cubin)
            res = subprocess.run(["cuobjdump", "--dump-resource-usage", f.name], capture_output=True, text=True).stdout.splitlines()
            use = [res[k + 1].strip() for k, line in enumerate(res) if f"Function {kname}:" in line]
            hist = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "sass_loop_hist.py"), f.name, kname],
                                  capture_output=True, text=True).stdout
            os.unlink(f.name)
            out.write(f"==== {tag}: {model} {stepper} {obs} min_blocks_per_sm={m} {kw or ''}\n{use[0] if use else ''}\n{hist}\n")
    print(path)


if __name__ == "__main__":
    main()
