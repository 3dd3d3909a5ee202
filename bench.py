#!/usr/bin/env python
"""bench.py — headline benchmark of the ensemble-ODE hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C5]

Workload (default C2 = BASELINE.json configs[1]): Lorenz system, `features` kernel, dopri5,
observer `basic`, 2^20 parameter sets PER GPU (weak scaling: the global r-grid is N*2^20 points,
rank g integrates the interleaved shard {g, g+N, g+2N, ...}, i.e. every GPU sees the same cost distribution), double precision, t in [0,100],
dt0 = 0.01, dtmax = 1, abstol = reltol = 1e-6.  One "step" = one pass of the hot
path (initializeObserver + features kernels) over the rank's ensemble.

metric  = accepted ODE instance-steps per second, whole job (sum over ranks / max-over-ranks time)
value   : inputs resident in HBM, kernel time by CUDA events on the launch stream
e2e     : the same metric through the public C-ABI call sequence with HOST buffers — per step
          H2D of x0/pars/dt from pinned memory, the kernels, D2H of the feature matrix
roofline: FP64 FMA pipe. achieved = 283 algorithmic flop per accepted Lorenz-dopri5 step
          (clode_b200/flops.py) x steps / kernel time; peak = DFMA micro-benchmark measured
          on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry), nominal quoted beside it.
cpu_baseline / --impl reference: the reference's own kernel sources compiled as host C
          (oracle/_ref, OpenMP over all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

N_PER_GPU = 1 << 20
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: SMs x FP64 lanes x 2 x max clock


# ------------------------------------------------------------------------------------------------
def workload(name: str, n_total: int, index):
    """inputs of the instances `index` of a global ensemble of n_total (variable-major, float64)"""
    from clode_b200.flops import flops_per_step

    idx = np.asarray(index, dtype=np.int64)
    n = idx.size
    frac = idx.astype(np.float64) / max(n_total - 1, 1)
    if name in ("C2", "C2l"):
        # BASELINE.json configs[1] names both observers: C2 = basic (the headline), C2l = localmax (26 features)
        w = dict(model="lorenz63", stepper="dopri5", observer="basic" if name == "C2" else "localmax", kind="features", tspan=(0.0, 100.0),
                 solver=dict(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, max_steps=10000000),
                 observer_params=dict(max_event_count=10000),
                 pars=np.concatenate([0.5 + 59.5 * frac, np.full(n, 10.0), np.full(n, 8.0 / 3.0)]),
                 x0=np.ones(3 * n), desc=f"{name}: Lorenz features, dopri5, observer {'basic' if name == 'C2' else 'localmax'}, 2^20 parameter sets per GPU, f64")
    elif name == "C3":
        # 1024 x 1024 (gcal x gbk) grid, flattened row-major; bs23 + thresh2 (two-pass)
        side = int(round(n_total ** 0.5))
        gcal = 0.5 + 3.5 * (idx // side) / max(side - 1, 1)
        gbk = 2.0 * (idx % side) / max(side - 1, 1)
        w = dict(model="lactotroph", stepper="bs23", observer="thresh2", kind="features", tspan=(0.0, 10000.0),
                 solver=dict(dt=0.1, dtmax=100.0, abstol=1e-6, reltol=1e-4, max_steps=10000000),
                 observer_params=dict(max_event_count=100000, x_up_threshold=0.3, x_down_threshold=0.2),
                 pars=np.concatenate([gcal, np.full(n, 3.0), gbk]),
                 x0=np.concatenate([np.full(n, -60.0), np.zeros(n), np.zeros(n), np.full(n, 0.1)]),
                 desc="C3: lactotroph thresh2 features, bs23, 1024x1024 grid per GPU, f64")
    elif name == "C4":
        # lactotroph + current noise, Euler-Maruyama, identical parameters, per-instance RNG streams
        w = dict(model="lactotroph_noise", stepper="seuler", observer="basicall", kind="features", tspan=(0.0, 100.0),
                 solver=dict(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-4, max_steps=10000000),
                 observer_params=dict(),
                 pars=np.concatenate([np.full(n, 1.5), np.full(n, 3.0), np.full(n, 1.0), np.full(n, 1.0)]),
                 x0=np.concatenate([np.full(n, -60.0), np.zeros(n), np.zeros(n), np.full(n, 0.1)]),
                 desc="C4: lactotroph_noise stochastic Euler features (basicall), per-instance RNG streams, f64")
    elif name in ("C5", "C5e"):
        # Chay-Keizer trajectories: 512 x 512 (gca x kpmca) grid per GPU, 2000 stored points, nout = 1
        side = int(round(n_total ** 0.5))
        gca = 550.0 + 500.0 * (idx // side) / max(side - 1, 1)
        kpmca = 0.095 + 0.06 * (idx % side) / max(side - 1, 1)
        stepper = "rk4" if name == "C5" else "euler"
        w = dict(model="chay_keizer", stepper=stepper, observer="basic", kind="trajectory", tspan=(0.0, 1000.0),
                 solver=dict(dt=0.5 if name == "C5" else 0.05, dtmax=1.0, abstol=1e-6, reltol=1e-4, max_steps=10000000,
                             max_store=2000, nout=1),
                 observer_params=dict(),
                 pars=np.concatenate([gca, np.full(n, 750.0), kpmca]),
                 x0=np.concatenate([np.full(n, -50.0), np.full(n, 0.01), np.full(n, 0.12)]),
                 desc=f"C5: Chay-Keizer trajectory, {stepper}, 2000 stored points x 2^18 instances per GPU, nout=1, f64")
    else:
        raise SystemExit(f"unknown workload {name}")
    w["flops_per_step"] = flops_per_step(w["stepper"], w["model"])
    return w


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference(wname: str, n_total: int, steps: int, warmup: int, stride: int = 4):
    """the reference's own kernels on the host cores (oracle/_ref; falls back to the C port)"""
    from oracle import ref, restate
    from oracle.common import Config, Observer, Solver, seed_states

    w = workload(wname, n_total, np.arange(0, N_PER_GPU))
    cfg = Config(w["model"], w["stepper"], w["observer"], contract="fast")
    if os.path.exists(ref.so_path(cfg)) or ref.reference_available():
        lib, kind = ref.RefLib(cfg), "reference"
    else:
        lib, kind = restate.OracleLib(cfg), "port"
    nv = lib.n_var
    sel = np.arange(0, N_PER_GPU, stride)
    n = sel.size
    x0 = w["x0"].reshape(nv, -1)[:, sel].ravel()
    pars = w["pars"].reshape(lib.n_par, -1)[:, sel].ravel()
    sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
    cores = os.cpu_count() or 1
    step_row = {"basic": 5}.get(w["observer"], lib.n_feat - (1 if w["observer"] in ("basicall", "localmax") else 4))
    times, total = [], 0
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        r = lib.features(w["tspan"], x0, pars, sp, op, np.full(n, sp.dt), seed_states(1, n), nthreads=cores)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
            total += int(r["F"].reshape(lib.n_feat, n)[step_row].sum())
    value = total / sum(times)
    sample = f"every {stride}th instance of the N=1 workload ({n} instances), {steps} passes, OpenMP {cores} threads, gcc -O3 -march=x86-64-v3"
    return dict(value=value, unit="instance-steps/s", cores=cores, kind=kind, sample=sample), sum(times) / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = cpu_reference(args.workload, N_PER_GPU * args.gpus, args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": "ODE instance-steps/sec (dopri5, 1M-param sweep)", "value": base["value"],
            "unit": "instance-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload(args.workload, N_PER_GPU, np.arange(1))["desc"],
                                            "note": "CPU arm runs a bounded sample; the metric is per instance-step"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_ours(args):
    import torch

    from clode_b200 import _rt, build
    from clode_b200.models import MODELS, rhs_source

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    build.build_runtime()
    default_n = {"C5": 1 << 18, "C5e": 1 << 18, "C4": 1 << 22}.get(args.workload, N_PER_GPU)
    n = default_n if args.npts <= 0 else args.npts
    n_total = n * world
    from clode_b200 import sharding
    # cost-balanced interleaved shards of ONE global grid: rank g owns instances g, g+N, g+2N, ...
    index = sharding.interleaved(n_total, world, rank)
    w = workload(args.workload, n_total, index)
    nv, npar, na, nw = MODELS[w["model"]]
    is_traj = w["kind"] == "trajectory"
    prog = _rt.Program(rhs_source(w["model"]), w["stepper"], nv, npar, na, nw, observer=w["observer"],
                       kernels=_rt.KERNEL_TRAJECTORY if is_traj else _rt.KERNEL_FEATURES,
                       work_queue=bool(args.work_queue), block_size=args.block, min_blocks_per_sm=args.min_blocks,
                       staged_trajectory=bool(args.staged), observer_in_shared=bool(args.obs_smem),
                       single_precision=bool(args.single), ieee_constant_division=bool(args.ieee_div),
                       library_exp=bool(args.library_exp), bit_exact=bool(args.bit_exact))
    sim = _rt.Sim(prog, device=local)
    sim.set_solver_params(**w["solver"])
    sim.set_observer_params(**w["observer_params"])
    sim.set_tspan(*w["tspan"])

    # host inputs in pinned memory (numpy views of pinned torch tensors)
    def pinned(a):
        t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[:] = a
        return t, v
    if args.shuffle:
        perm = np.random.default_rng(args.shuffle).permutation(n)
        w["pars"] = w["pars"].reshape(npar, n)[:, perm].ravel()
        w["x0"] = w["x0"].reshape(nv, n)[:, perm].ravel()
        w["desc"] += ", grid order shuffled"
    keep_x0, x0 = pinned(w["x0"])
    keep_p, pars = pinned(w["pars"])
    keep_dt, dt0 = pinned(np.full(n, w["solver"]["dt"]))
    sim.set_problem(x0, pars)
    sim.set_rng_state(sharding.seed_states_for(1, n_total, index))  # global seeding rule (CLODE.cpp:447-453)
    nfeat = sim.n_features() if not is_traj else nv  # trajectory: the gathered / fetched result is xf
    step_row = {"basic": 5}.get(w["observer"], nfeat - (1 if w["observer"] in ("basicall", "localmax") else 4))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_features():
        """the one exchange step of the path: features of every shard to rank 0 over NVLink (NCCL)"""
        if not dist:
            return None
        ptr, nbytes, _ = sim.device_buffer(_rt.BUF_XF if is_traj else _rt.BUF_F)
        local_f = torch.as_tensor(_CudaArray(ptr, nbytes // 8, "<f8"), device=f"cuda:{local}")
        return sharding.gather_interleaved(local_f, nfeat, n_total)

    def hot_step():
        sim.set_dt(dt0)          # per-instance dt persists across calls (reference semantics): reset it
        if is_traj:
            sim.trajectory()     # trajectory kernel, synchronous
        else:
            sim.features(1)      # initializeObserver + features kernels, synchronous
        ms = sim.last_kernel_ms()
        gather_features()
        return ms

    # ---- warm-up, then K timed steps with inputs resident in HBM ------------------------------
    first_call_ms = None
    for k in range(max(args.warmup, 3)):
        ms = hot_step()
        if k == 0:
            first_call_ms = ms   # no history yet: blocks in forward order (see config.block_order)
    launches0 = sim.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t_wall = time.perf_counter()
    kernel_ms = [hot_step() for _ in range(args.steps)]
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - launches0
    steps_per_pass = int(sim.get_steps().astype(np.int64).sum())
    dev_ms = sum(kernel_ms)
    # the same step with the history-based block order switched off (reported next to the headline, not part of it)
    order_env = os.environ.get("CLODE_BLOCK_ORDER")
    os.environ["CLODE_BLOCK_ORDER"] = "forward"
    forward_ms = min(hot_step() for _ in range(3))
    if order_env is None:
        del os.environ["CLODE_BLOCK_ORDER"]
    else:
        os.environ["CLODE_BLOCK_ORDER"] = order_env
    hot_step()  # restore the history for the end-to-end leg

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    keep_out, out_host = pinned(np.zeros(n * nfeat))  # pinned landing buffer for the per-step result read-back
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.set_x0(x0)
        sim.set_pars(pars)
        sim.set_dt(dt0)
        if is_traj:
            # the e2e result read back is the final state; the 29 GB trajectory itself stays on the device
            # (fetching it is a separate API call, CLODEtrajectory::getX)
            sim.trajectory()
            F = sim.get_xf(out_host)
        else:
            sim.features(1)
            F = sim.get_f(out_host)
        gather_features()
    barrier()
    e2e_s = time.perf_counter() - t0
    if not is_traj:
        assert int(F.reshape(nfeat, n)[step_row].sum()) == steps_per_pass
    h2d = 8 * n * (nv + npar + 1)
    d2h = 8 * n * nfeat

    # ---- trajectory only: bringing the stored points themselves to the host -------
    fetch = None
    if is_traj and args.stream_rows > 0:
        rows = w["solver"]["max_store"]
        sizes = dict(t=rows * n, x=rows * n * nv, dx=rows * n * nv, aux=rows * n * na)
        host = {k: _rt.pinned_empty(v, device=local) for k, v in sizes.items() if v}
        which = dict(t=_rt.BUF_T, x=_rt.BUF_X, dx=_rt.BUF_DX, aux=_rt.BUF_AUX)

        def single_launch_then_copy():
            sim.set_dt(dt0)
            sim.trajectory()
            for k, a in host.items():
                sim.get(which[k], a.size, out=a)

        def streamed():
            sim.set_dt(dt0)
            sim.trajectory_stream(args.stream_rows, out=host, want=tuple(host))

        times = {}
        for name, fn in (("single_launch_then_copy", single_launch_then_copy), ("streamed", streamed)):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            barrier()
            times[name] = (time.perf_counter() - t0) / args.steps
            check = float(host["x"].reshape(rows, nv, n)[rows // 2].sum())
            times[name + "_checksum"] = check
        nbytes = 8 * sum(sizes.values())
        fetch = {"rows": rows, "host_bytes": nbytes, "chunk_rows": args.stream_rows,
                 "single_launch_then_copy_s": times["single_launch_then_copy"], "streamed_s": times["streamed"],
                 "single_launch_then_copy_GBps": nbytes / times["single_launch_then_copy"] / 1e9,
                 "streamed_GBps": nbytes / times["streamed"] / 1e9,
                 "device_bytes_single_launch": 8 * (rows + 1) * n * (1 + 2 * nv + na),
                 "device_bytes_streamed": 2 * 8 * min(args.stream_rows, rows + 1) * n * (1 + 2 * nv + na),
                 "same_result": times["single_launch_then_copy_checksum"] == times["streamed_checksum"]}

    # ---- reduce over ranks -----------------------------------------------------------------------
    stats = torch.tensor([dev_ms, wall_ms, e2e_s, float(steps_per_pass), float(launches)], dtype=torch.float64, device="cuda")
    if dist:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, e2e_s = mx[0].item(), mx[1].item(), mx[2].item()
        total_steps_per_pass, launches = int(sm[3].item()), int(sm[4].item())
    else:
        total_steps_per_pass = steps_per_pass
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    value = total_steps_per_pass * args.steps / (dev_ms * 1e-3)
    peak_tf, _ = _rt.measure_fp64_peak(local, 5)
    achieved_tf = w["flops_per_step"] * steps_per_pass * args.steps / (sum(kernel_ms) * 1e-3) / 1e12
    info = sim.kernel_info(_rt.KERNEL_TRAJECTORY if is_traj else _rt.KERNEL_FEATURES)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    # (profiles/r01e_*_summary.txt; C3: the features kernel, the dominant launch of the pair — its writes are spill
    # lines evicted from L2, not results); None for workloads that were not captured
    ncu_traffic = {"C2": 141.124864e6 + 121.296640e6, "C3": 760.737280e6 + 2.706634e9,
                   "C5": 21.410816e6 + 29.337310e9, "C5e": 20.577280e6 + 29.331946e9}
    traffic = ncu_traffic.get(args.workload) if n == default_n else None
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "peak_source": "DFMA micro-benchmark measured on this GPU in this run (clode_measure_fp64_peak)",
                "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved_tf / NOMINAL_FP64_TFLOPS,
                "flops_per_step": w["flops_per_step"]}
    if is_traj:
        from clode_b200.flops import trajectory_bytes_per_point
        stored = int(sim.get_trajectory_counts().astype(np.int64).sum()) + n  # + the initial point of every instance
        per_point = trajectory_bytes_per_point(w["model"], na)
        gbs = per_point * stored * args.steps / (sum(kernel_ms) * 1e-3) / 1e9
        peaks_file = os.path.join(REPO, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                    "bytes_per_stored_point": per_point, "stored_points": stored,
                    "fp64": {"achieved": achieved_tf, "peak": peak_tf, "frac": achieved_tf / peak_tf}}
    base, _ = cpu_reference(args.workload, n_total, 1, 1) if not (args.no_cpu_baseline or is_traj) else (None, None)
    line = {
        "metric": "ODE instance-steps/sec (dopri5, 1M-param sweep)", "value": value, "unit": "instance-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.single else "f64",
        "data": "synthetic",
        "config": {"workload": w["desc"] if not args.single else w["desc"].replace("f64", "f32"), "instances_per_gpu": n, "accepted_steps_per_pass": total_steps_per_pass,
                   "l2": "inputs+outputs per pass exceed the 126 MB L2; the kernel is FP64-pipe bound, not memory bound",
                   "kernel": info, "work_queue": bool(args.work_queue), "wall_ms_per_step": wall_ms / args.steps,
                   "block_order": os.environ.get("CLODE_BLOCK_ORDER", "auto") + ": each launch walks the ensemble forwards, or backwards when "
                                  "the previous launch on the same ensemble spent more accepted steps in the upper half of the index range "
                                  "(scheduling only; all work is redone every step)",
                   "first_call_ms": first_call_ms, "forward_order_ms_per_step": forward_ms},
        "clocks": clocks,
        "e2e": {"value": total_steps_per_pass * args.steps / e2e_s, "unit": "instance-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": base,
    }
    if fetch:
        line["trajectory_fetch"] = fetch
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--npts", type=int, default=0, help="instances per GPU (default 2^20)")
    ap.add_argument("--work-queue", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--min-blocks", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stream-rows", type=int, default=0,
                    help="trajectory workloads: also time fetching all stored points, single launch + copy vs streamed in chunks of this many rows")
    ap.add_argument("--staged", type=int, default=0, help="trajectory: shared-memory staged TMA bulk stores")
    ap.add_argument("--obs-smem", type=int, default=0, help="features: observer state in shared memory")
    ap.add_argument("--single", type=int, default=0, help="single precision (the reference's Python default); not the headline")
    ap.add_argument("--ieee-div", type=int, default=0, help="leave `x / literal` to ptxas' generic division (A/B of the PTX pass)")
    ap.add_argument("--library-exp", type=int, default=0, help="CUDA's exp instead of device/fast_exp.cuh (A/B)")
    ap.add_argument("--bit-exact", type=int, default=0, help="bit-exact tier (no FMA contraction, portable math): the tier with identical step counts")
    ap.add_argument("--shuffle", type=int, default=0, help="randomly permute the parameter grid (heterogeneous warps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
