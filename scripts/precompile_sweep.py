#!/usr/bin/env python
"""Fill clode_b200/_cubin_cache with the programs a sweep spec (scripts/sweeps/*.spec) will build on the GPU box, so that
the box's GPU-minutes go into running, not into NVRTC.  No GPU needed (NVRTC + ptxas cross-compile for sm_100a).

  python scripts/precompile_sweep.py scripts/sweeps/<name>.spec

Each spec line is "label | ENV=VALUE ... | bench.py arguments" (scripts/gpu_sweep.sh).  For every line the occupancy
candidates the runtime tries first (clode_sim_build: 5, 4, 3 blocks of 128 threads; the one given by --min-blocks) are
compiled under the line's environment, with and without the observer extents in shared memory where the runtime decides
that from the spill size on the box."""
import argparse
import dataclasses
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    spec = sys.argv[1]
    from clode_b200 import _rt, build
    from clode_b200.models import MODELS, rhs_source
    import bench

    build.build_runtime()
    ap = argparse.ArgumentParser()
    for flag, kind, default in (("--workload", str, "C2"), ("--block", int, 0), ("--min-blocks", int, 0), ("--work-queue", int, 0),
                                ("--staged", int, 0), ("--obs-smem", int, 0), ("--single", int, 0), ("--ieee-div", int, 0),
                                ("--library-exp", int, 0), ("--bit-exact", int, 0)):
        ap.add_argument(flag, type=kind, default=default)
    done = 0
    for line in open(spec):
        parts = [p.strip() for p in line.split("|")]
        if len(parts) < 3 or not parts[0]:
            continue
        envs = dict(tok.split("=", 1) for tok in parts[1].split() if "=" in tok)
        args, _ = ap.parse_known_args(parts[2].split())
        name = args.workload
        w = bench.workload("C2" if name == "C2t" else name, 256, __import__("numpy").arange(256))
        nv, npar, na, nw = MODELS[w["model"]]
        is_traj = w["kind"] == "trajectory"
        kernels = _rt.KERNEL_TRAJECTORY if is_traj else (_rt.KERNEL_TRANSIENT if name in ("C2t", "C1") else _rt.KERNEL_FEATURES)
        prog = _rt.Program(rhs_source(w["model"]), w["stepper"], nv, npar, na, nw, observer=w["observer"], kernels=kernels,
                           work_queue=bool(args.work_queue), block_size=args.block, min_blocks_per_sm=args.min_blocks,
                           staged_trajectory=bool(args.staged), observer_in_shared=bool(args.obs_smem),
                           single_precision=bool(args.single), ieee_constant_division=bool(args.ieee_div),
                           library_exp=bool(args.library_exp), bit_exact=bool(args.bit_exact))
        saved = {k: os.environ.get(k) for k in envs}
        os.environ.update(envs)
        try:
            per_sm = 640 // (args.block or 128)
            blocks = [args.min_blocks] if args.min_blocks else [m for m in (per_sm, per_sm - 1, per_sm - 2) if m >= 1]
            variants = [None]
            if w["observer"] != "basic" and kernels == _rt.KERNEL_FEATURES and "CLODE_EXT_SMEM" not in envs:
                variants.append("1")
            for ext in variants:
                if ext:
                    os.environ["CLODE_EXT_SMEM"] = ext
                for m in blocks:
                    cubin, _ = _rt.compile_program(dataclasses.replace(prog, min_blocks_per_sm=m))
                    assert cubin[:4] == b"\x7fELF"
                    done += 1
                if ext:
                    del os.environ["CLODE_EXT_SMEM"]
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        print(f"{parts[0]}: ok", flush=True)
    print(f"{done} programs in the cache")


if __name__ == "__main__":
    main()
