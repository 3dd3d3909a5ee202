#!/usr/bin/env python
"""bench.py — headline benchmark of the ensemble-ODE hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C4|C5] [--quick 1]

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
          (oracle/_ref, OpenMP over all host cores) on a bounded sample of the same workload: every (4N)-th instance
          of the SAME global grid the N-GPU job integrates.

Beside the headline the JSON line carries (skipped with --quick 1):
  tiers        : the bit-exact tier (no FMA contraction, portable math; identical accepted-step counts to the oracle)
                 timed on the same workload next to the production tier
  parity       : after the timed region, a sample of instances of BOTH tiers against the CPU oracle
  strong       : the same 2^20-instance sweep split over the N GPUs (strong scaling; the headline is weak scaling)
  e2e_frontend : the metric through the user-facing Python API (FeatureSimulator.set_ensemble -> features() ->
                 ObserverOutput) with ordinary numpy arrays; at N > 1 one process drives all N GPUs (device_ids=...)
  config.secondary : one-line results of the other BASELINE configs (C3, C4, C5, C2 localmax, the transient kernel)
  small_n      : launch + kernel time for N = 2^0 .. 2^17 instances (examples/dump_device_performance.py:36-105)
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
    elif name == "C1":
        # BASELINE.json configs[0]: Van der Pol transient, rk4, 4096-point mu grid (test_vdp.py's path); the transient kernel
        w = dict(model="vanderpol", stepper="rk4", observer="basic", kind="features", tspan=(0.0, 100.0),
                 solver=dict(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=10000000),
                 observer_params=dict(), pars=0.1 + 9.9 * frac, x0=np.ones(2 * n),
                 desc="C1: Van der Pol TRANSIENT kernel, rk4 dt=0.01, t in [0,100], 4096-point mu grid per GPU, f64")
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
    elif name in ("C5", "C5e", "C5d"):
        # Chay-Keizer trajectories: 512 x 512 (gca x kpmca) grid per GPU, 2000 stored points, nout = 1
        side = int(round(n_total ** 0.5))
        gca = 550.0 + 500.0 * (idx // side) / max(side - 1, 1)
        kpmca = 0.095 + 0.06 * (idx % side) / max(side - 1, 1)
        stepper = {"C5": "rk4", "C5e": "euler", "C5d": "dopri5"}[name]  # BASELINE configs[4]: rk4 (and dopri5); C5e = store-bound variant
        w = dict(model="chay_keizer", stepper=stepper, observer="basic", kind="trajectory", tspan=(0.0, 1000.0),
                 solver=dict(dt=0.5 if name == "C5" else 0.05, dtmax=1.0, abstol=1e-6, reltol=1e-4, max_steps=10000000,
                             max_store=2000, nout=1),
                 observer_params=dict(),
                 pars=np.concatenate([gca, np.full(n, 750.0), kpmca]),
                 x0=np.concatenate([np.full(n, -50.0), np.full(n, 0.01), np.full(n, 0.12)]),
                 desc=f"{name}: Chay-Keizer trajectory, {stepper}, 2000 stored points x 2^18 instances per GPU, nout=1, f64")
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


STEP_ROW = {"basic": 5}


def step_row(observer: str, n_feat: int) -> int:
    return STEP_ROW.get(observer, n_feat - (1 if observer in ("basicall", "localmax") else 4))


def cpu_reference(wname: str, n_gpus: int, steps: int, warmup: int, sample: int = 1 << 18):
    """the reference's own kernels on the host cores (oracle/_ref; falls back to the C port) on a bounded sample of
    the GLOBAL grid the N-GPU job integrates: every (n_total / sample)-th instance"""
    from oracle import ref, restate
    from oracle.common import Config, Observer, Solver
    from clode_b200 import sharding
    from clode_b200.models import MODELS

    per_gpu = {"C5": 1 << 18, "C5e": 1 << 18, "C5d": 1 << 18, "C4": 1 << 22}.get(wname, N_PER_GPU)
    n_total = per_gpu * n_gpus
    sample = min(sample, n_total)
    stride = n_total // sample
    sel = np.arange(0, n_total, stride)[:sample]
    w = workload(wname, n_total, sel)
    cfg = Config(w["model"], w["stepper"], w["observer"], contract="fast")
    if os.path.exists(ref.so_path(cfg)) or ref.reference_available():
        lib, kind = ref.RefLib(cfg), "reference"
    else:
        lib, kind = restate.OracleLib(cfg), "port"
    n = sel.size
    sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
    cores = os.cpu_count() or 1
    row = step_row(w["observer"], lib.n_feat)
    times, total = [], 0
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        r = lib.features(w["tspan"], w["x0"], w["pars"], sp, op, np.full(n, sp.dt), sharding.seed_states_for(1, n_total, sel), nthreads=cores)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
            total += int(r["F"].reshape(lib.n_feat, n)[row].sum())
    value = total / sum(times)
    desc = (f"every {stride}th instance of the global grid of the {n_gpus}-GPU job ({n} of {n_total} instances, the same parameter "
            f"range), {steps} passes, OpenMP {cores} threads, gcc -O3 -march=x86-64-v3")
    return dict(value=value, unit="instance-steps/s", cores=cores, kind=kind, sample=desc), sum(times) / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm (it uses all host cores); the other ranks exit without work
    base, ms = cpu_reference(args.workload, args.gpus, args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": "ODE instance-steps/sec (dopri5, 1M-param sweep)", "value": base["value"],
            "unit": "instance-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload(args.workload, N_PER_GPU, np.arange(1))["desc"],
                                            "note": "CPU arm, rank 0 only, all host cores; bounded sample of the job's global grid; the metric is per instance-step"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class Job:
    """one workload on this rank's GPU through the C ABI (clode_b200._rt.Sim)"""

    def __init__(self, args, name, n, world, rank, local, bit_exact=False, kernels=None, shuffle=0):
        from clode_b200 import _rt, sharding
        from clode_b200.models import MODELS, rhs_source

        self.rt, self.name, self.n, self.world, self.rank = _rt, name, n, world, rank
        self.n_total = n * world
        self.index = sharding.interleaved(self.n_total, world, rank)  # rank g owns instances g, g+N, g+2N, ... of ONE global grid
        self.transient = name in ("C2t", "C1")  # the transient kernel (north_star: FMA-pipe utilisation of features AND transient)
        w = self.w = workload("C2" if name == "C2t" else name, self.n_total, self.index)
        if name == "C2t":
            w["desc"] = "C2t: Lorenz TRANSIENT kernel, dopri5, 2^20 parameter sets per GPU, f64"
        nv, npar, na, nw = self.dims = MODELS[w["model"]]
        self.is_traj = w["kind"] == "trajectory"
        if kernels is None:
            kernels = _rt.KERNEL_TRAJECTORY if self.is_traj else (_rt.KERNEL_TRANSIENT if self.transient else _rt.KERNEL_FEATURES)
        self.kernel_id = kernels
        prog = _rt.Program(rhs_source(w["model"]), w["stepper"], nv, npar, na, nw, observer=w["observer"], kernels=kernels,
                           work_queue=bool(args.work_queue), block_size=args.block, min_blocks_per_sm=args.min_blocks,
                           staged_trajectory=bool(args.staged), observer_in_shared=bool(args.obs_smem),
                           single_precision=bool(args.single), ieee_constant_division=bool(args.ieee_div),
                           library_exp=bool(args.library_exp), bit_exact=bool(bit_exact))
        sim = self.sim = _rt.Sim(prog, device=local)
        sim.set_solver_params(**w["solver"])
        sim.set_observer_params(**w["observer_params"])
        sim.set_tspan(*w["tspan"])
        if shuffle:
            perm = np.random.default_rng(shuffle).permutation(n)
            w["pars"] = w["pars"].reshape(npar, n)[:, perm].ravel()
            w["x0"] = w["x0"].reshape(nv, n)[:, perm].ravel()
            w["desc"] += ", grid order shuffled"
        self.x0 = _rt.pinned_empty(nv * n, device=local); self.x0[:] = w["x0"]
        self.pars = _rt.pinned_empty(npar * n, device=local); self.pars[:] = w["pars"]
        self.dt0 = _rt.pinned_empty(n, device=local); self.dt0[:] = w["solver"]["dt"]
        sim.set_problem(self.x0, self.pars)
        sim.set_rng_state(sharding.seed_states_for(1, self.n_total, self.index))  # global seeding rule (CLODE.cpp:447-453)
        self.n_feat = nv if (self.is_traj or self.transient) else sim.n_features()
        self.step_row = step_row(w["observer"], self.n_feat)

    def step(self):
        """one pass of the hot path over this rank's ensemble, inputs resident in HBM; returns device ms"""
        s = self.sim
        s.set_dt(self.dt0)  # per-instance dt persists across calls (reference semantics): reset it
        if self.is_traj:
            s.trajectory()
        elif self.transient:
            s.transient()
        else:
            s.features(1)  # initializeObserver + features kernels (pilot + sorted rounds for the adaptive steppers)
        return s.last_kernel_ms()

    def steps_per_pass(self):
        return int(self.sim.get_steps().astype(np.int64).sum())

    def info(self):
        return self.sim.kernel_info(self.kernel_id)

    def close(self):
        self.sim.close()


def run_ours(args):
    import torch

    from clode_b200 import _rt, build, sharding
    from clode_b200.models import MODELS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build_runtime()

    default_n = {"C1": 4096, "C5": 1 << 18, "C5e": 1 << 18, "C5d": 1 << 18, "C4": 1 << 22}.get(args.workload, N_PER_GPU)
    n = default_n if args.npts <= 0 else args.npts
    job = Job(args, args.workload, n, world, rank, local, bit_exact=bool(args.bit_exact), shuffle=args.shuffle)
    sim, w = job.sim, job.w
    nv, npar, na, nw = job.dims
    n_total, nfeat, is_traj = job.n_total, job.n_feat, job.is_traj

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if not dist:
            return list(values), list(values)
        mx, sm = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        return mx.tolist(), sm.tolist()

    # the one exchange step of the path: results of every shard to rank 0 over NVLink (NCCL gather), issued asynchronously
    # from a staging copy so that it overlaps the next pass (clode_b200/sharding.py PipelinedGather)
    gatherer = sharding.PipelinedGather(nfeat, n_total) if dist else None

    def gather_async(k):
        if not dist:
            return
        ptr, nbytes, _ = sim.device_buffer(_rt.BUF_XF if (is_traj or job.transient) else _rt.BUF_F)
        gatherer.submit(torch.as_tensor(_CudaArray(ptr, nbytes // 8, "<f8"), device=f"cuda:{local}"))

    def gather_drain():
        out = gatherer.drain() if dist else None
        torch.cuda.synchronize()
        return out

    # ---- warm-up, then K timed steps with inputs resident in HBM ------------------------------
    W = max(args.warmup, 3)
    first_call_ms = None
    for k in range(W):
        ms = job.step()
        gather_async(k)
        if k == 0:
            first_call_ms = ms
    gather_drain()
    launches0 = sim.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t_wall = time.perf_counter()
    kernel_ms = []
    for k in range(args.steps):
        kernel_ms.append(job.step())
        gather_async(k)
    gather_drain()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - launches0
    steps_per_pass = job.steps_per_pass()
    dev_ms = sum(kernel_ms)

    # ---- end to end through the C ABI with HOST buffers: H2D of x0 / pars / dt, the kernels, D2H of the result,
    #      the NVLink gather -------------------------------------------------------------------------------------
    out_host = _rt.pinned_empty(n * nfeat, device=local)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sim.set_x0(job.x0)
        sim.set_pars(job.pars)
        sim.set_dt(job.dt0)
        if is_traj:
            sim.trajectory()       # the e2e result read back is the final state; fetching the 29 GB trajectory itself
            F = sim.get_xf(out_host)  # is a separate API call (CLODEtrajectory::getX), see --stream-rows
        elif job.transient:
            sim.transient()
            F = sim.get_xf(out_host)
        else:
            sim.features(1)
            F = sim.get_f(out_host)
        gather_async(k)
    gathered = gather_drain()
    barrier()
    e2e_s = time.perf_counter() - t0
    if not is_traj and not job.transient:
        assert int(F.reshape(nfeat, n)[job.step_row].sum()) == steps_per_pass
        if gathered is not None:  # rank 0: the gathered global matrix carries every rank's accepted steps
            gathered_steps = int(gathered.view(nfeat, n_total)[job.step_row].sum().item())
    h2d = 8 * n * (nv + npar + 1)
    d2h = 8 * n * nfeat

    # ---- trajectory only: bringing the stored points themselves to the host -------
    fetch = None
    if is_traj and args.stream_rows > 0:
        fetch = trajectory_fetch(args, job, barrier)

    (dev_ms, wall_ms, e2e_s), (_, _, _) = reduce_max_sum([dev_ms, wall_ms, e2e_s])
    _, (total_steps_per_pass, launches) = reduce_max_sum([float(steps_per_pass), float(launches)])
    total_steps_per_pass, launches = int(total_steps_per_pass), int(launches)
    if dist and rank == 0 and not is_traj and not job.transient:
        assert gathered_steps == total_steps_per_pass, (gathered_steps, total_steps_per_pass)

    extras = {}
    if not args.quick and not is_traj and not job.transient:
        extras = run_extras(args, job, world, rank, local, dist, barrier, reduce_max_sum)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    value = total_steps_per_pass * args.steps / (dev_ms * 1e-3)
    peak_tf, _ = _rt.measure_fp64_peak(local, 5)
    achieved_tf = w["flops_per_step"] * steps_per_pass * args.steps / (sum(kernel_ms) * 1e-3) / 1e12
    info = job.info()
    roofline = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": None,
                "traffic_note": "not measured in this run (the kernel is FP64-pipe bound); the ncu captures under profiles/ give dram bytes per launch",
                "peak_source": "DFMA micro-benchmark measured on this GPU in this run (clode_measure_fp64_peak)",
                "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved_tf / NOMINAL_FP64_TFLOPS,
                "flops_per_step": w["flops_per_step"],
                "launches_per_step": launches / max(args.steps, 1) / world,
                "note": "achieved = algorithmic flop of one pass / device time of ALL launches of the pass (pilot, sorted rounds and the three scheduler kernels between them), CUDA events on the launch stream"}
    if is_traj:
        from clode_b200.flops import trajectory_bytes_per_point
        stored = int(sim.get_trajectory_counts().astype(np.int64).sum()) + n  # + the initial point of every instance
        per_point = trajectory_bytes_per_point(w["model"], na)
        gbs = per_point * stored * args.steps / (sum(kernel_ms) * 1e-3) / 1e9
        peaks_file = os.path.join(REPO, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                    "bytes_per_stored_point": per_point, "stored_points": stored,
                    "fp64": {"achieved": achieved_tf, "peak": peak_tf, "frac": achieved_tf / peak_tf}}
    base = None
    if not (args.no_cpu_baseline or is_traj or world > 1):
        base, _ = cpu_reference(args.workload, world, 1, 1)
    sched = os.environ.get("CLODE_SCHED", "default (pilot 64 attempts for dopri5 / 256 for bs23, 12 sorted rounds, budget = 0.35 x dearest predicted remaining cost, >= pilot)")
    build_knobs = {k: os.environ.get(k, "default") for k in ("CLODE_BRANCHLESS", "CLODE_FAST_POLAR", "CLODE_EXT_SMEM", "CLODE_EXT_BATCH", "CLODE_IMM_HOIST")}
    line = {
        "metric": "ODE instance-steps/sec (dopri5, 1M-param sweep)", "value": value, "unit": "instance-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.single else "f64",
        "data": "synthetic",
        "config": {"workload": w["desc"] if not args.single else w["desc"].replace("f64", "f32"), "instances_per_gpu": n,
                   "tier": "bit-exact" if args.bit_exact else "production", "accepted_steps_per_pass": total_steps_per_pass,
                   "l2": "inputs+outputs per pass exceed the 126 MB L2; the kernel is FP64-pipe bound, not memory bound",
                   "kernel": info, "work_queue": bool(args.work_queue), "wall_ms_per_step": wall_ms / args.steps,
                   "scheduling": sched + ": every pass is pilot + cost-sorted rounds decided on the device from that pass alone — "
                                         "no history between passes, so the first call costs what every call costs",
                   "first_call_ms": first_call_ms,
                   "build": dict(build_knobs, note="production double defaults: branch-free exp / rcp / div, table-log polar method, "
                                                   "observer extents in shared memory when the features kernel spills (DESIGN.md section 3)")},
        "clocks": clocks,
        "e2e": {"value": total_steps_per_pass * args.steps / e2e_s, "unit": "instance-steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "C ABI (clode_sim_set_x0/_pars/_dt, clode_sim_features, clode_sim_get) with page-locked host buffers; "
                        "at N > 1 plus the asynchronous NCCL gather of the feature matrices to rank 0"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": base,
    }
    line.update(extras.get("top", {}))
    if extras.get("secondary"):
        line["config"]["secondary"] = extras["secondary"]
    if fetch:
        line["trajectory_fetch"] = fetch
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def trajectory_fetch(args, job, barrier):
    rt, sim, w = job.rt, job.sim, job.w
    nv, npar, na, nw = job.dims
    n, local = job.n, 0
    rows = w["solver"]["max_store"]
    sizes = dict(t=rows * n, x=rows * n * nv, dx=rows * n * nv, aux=rows * n * na)
    host = {k: rt.pinned_empty(v) for k, v in sizes.items() if v}
    which = dict(t=rt.BUF_T, x=rt.BUF_X, dx=rt.BUF_DX, aux=rt.BUF_AUX)

    def single_launch_then_copy():
        sim.set_dt(job.dt0)
        sim.trajectory()
        for k, a in host.items():
            sim.get(which[k], a.size, out=a)

    def streamed():
        sim.set_dt(job.dt0)
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
        times[name + "_checksum"] = float(host["x"].reshape(rows, nv, n)[rows // 2].sum())
    nbytes = 8 * sum(sizes.values())
    return {"rows": rows, "host_bytes": nbytes, "chunk_rows": args.stream_rows,
            "single_launch_then_copy_s": times["single_launch_then_copy"], "streamed_s": times["streamed"],
            "single_launch_then_copy_GBps": nbytes / times["single_launch_then_copy"] / 1e9,
            "streamed_GBps": nbytes / times["streamed"] / 1e9,
            "device_bytes_single_launch": 8 * (rows + 1) * n * (1 + 2 * nv + na),
            "device_bytes_streamed": 2 * 8 * min(args.stream_rows, rows + 1) * n * (1 + 2 * nv + na),
            "same_result": times["single_launch_then_copy_checksum"] == times["streamed_checksum"]}


# ------------------------------------------------------------------------------------------------
def timed(job, steps, warmup, barrier, reduce_max_sum):
    """(max-over-ranks device ms per step, whole-job steps per pass)"""
    for _ in range(warmup):
        job.step()
    barrier()
    ms = sum(job.step() for _ in range(steps))
    barrier()
    (ms,), _ = reduce_max_sum([ms])
    _, (total,) = reduce_max_sum([float(job.steps_per_pass())])
    return ms / steps, int(total)


def parity_sample(job, k=256):
    """a sample of this rank's instances of the last pass against the CPU oracle: production tier vs the reference
    arithmetic (libm, no contraction), bit-exact tier vs the portable-math oracle"""
    from clode_b200 import sharding
    from oracle import restate
    from oracle.common import Config, Observer, Solver

    w, n = job.w, job.n
    nv, npar, na, nw = job.dims
    bit_exact = job.sim.prog.bit_exact
    sub = np.sort(np.random.default_rng(5).choice(n, min(k, n), replace=False))
    F = job.sim.get_f().reshape(job.n_feat, n)[:, sub]
    lib = restate.OracleLib(Config(w["model"], w["stepper"], w["observer"], math="pm" if bit_exact else "libm"))
    sp, op = Solver(**w["solver"]), Observer(**w["observer_params"])
    o = lib.features(w["tspan"], sharding.take_rows(w["x0"], nv, n, sub), sharding.take_rows(w["pars"], npar, n, sub), sp, op,
                     np.full(sub.size, sp.dt), sharding.seed_states_for(1, job.n_total, job.index[sub]), nthreads=os.cpu_count() or 1)
    Fo = o["F"].reshape(job.n_feat, sub.size)
    row = job.step_row
    out = {"sampled_instances": int(sub.size), "oracle": "oracle/restate, " + ("portable math, no contraction" if bit_exact else "libm, no contraction"),
           "identical_accepted_step_counts": float((F[row] == Fo[row]).mean())}
    if bit_exact:
        out["bit_identical_features"] = bool(np.array_equal(F, Fo))
    else:
        scale = np.maximum(np.abs(Fo).max(axis=1, keepdims=True), 1e-300)
        if w["model"] == "lorenz63":
            r = w["pars"].reshape(npar, n)[0, sub]
            calm = r < 13.9
            out["lorenz_r_below_13.9"] = {"instances": int(calm.sum()), "identical_accepted_step_counts": float((F[row] == Fo[row])[calm].mean()),
                                          "max_feature_deviation_rel": float((np.abs(F - Fo) / scale)[:row, calm].max())}
            out["note"] = "r > 24.74 is chaotic (e^{0.9 t} error growth over t = 100): step counts of individual instances are not comparable there"
        out["max_rel_step_count_difference"] = float((np.abs(F[row] - Fo[row]) / np.maximum(Fo[row], 1)).max())
    return out


def run_extras(args, job, world, rank, local, dist, barrier, reduce_max_sum):
    """everything reported beside the headline; every rank takes part in the multi-GPU legs"""
    from clode_b200 import _rt

    top, secondary = {}, {}
    peak_tf = _rt.measure_fp64_peak(local, 3)[0] if rank == 0 else 1.0
    name = args.workload

    def line(j, ms, total):
        tf = j.w["flops_per_step"] * total / (ms * 1e-3) / 1e12 / world
        return {"workload": j.w["desc"], "ms_per_step": ms, "value": total / (ms * 1e-3), "unit": "instance-steps/s", "n_gpus": world,
                "fp64_tflops_per_gpu": tf, "frac_of_nominal_fp64": tf / NOMINAL_FP64_TFLOPS, "kernel": j.info()}

    # ---- parity of what was just timed, then the other tier on the same workload ------------------------------
    if rank == 0:
        top["parity"] = {("bit_exact_tier" if args.bit_exact else "production_tier"): parity_sample(job)}
    other = Job(args, name, job.n, world, rank, local, bit_exact=not args.bit_exact)
    ms, total = timed(other, 2, 1, barrier, reduce_max_sum)
    if rank == 0:
        top["parity"]["production_tier" if args.bit_exact else "bit_exact_tier"] = parity_sample(other)
        tf = other.w["flops_per_step"] * total / (ms * 1e-3) / 1e12 / world
        top["tiers"] = {"timed_as_headline": "bit-exact" if args.bit_exact else "production",
                        "other_tier": {"tier": "production" if args.bit_exact else "bit-exact", "ms_per_step": ms, "value": total / (ms * 1e-3),
                                       "unit": "instance-steps/s", "frac_of_nominal_fp64": tf / NOMINAL_FP64_TFLOPS, "kernel": other.info()},
                        "note": "bit-exact tier: --fmad=false + portable math, every output bit-identical to the CPU oracle (identical accepted-step "
                                "counts); production tier: FMA contraction + the instruction diet of DESIGN.md §3, within the tolerances of "
                                "tests/test_gpu_production_parity.py"}
    other.close()

    # ---- strong scaling: ONE 2^20-instance sweep split over the N GPUs ----------------------------------------------
    if world > 1:
        strong = Job(args, name, job.n // world, world, rank, local, bit_exact=bool(args.bit_exact))
        ms, total = timed(strong, max(args.steps // 2, 3), 2, barrier, reduce_max_sum)
        if rank == 0:
            top["strong"] = {"instances_total": strong.n_total, "instances_per_gpu": strong.n, "ms_per_step": ms, "value": total / (ms * 1e-3),
                             "unit": "instance-steps/s", "scaling": "strong"}
        strong.close()

    # ---- the other BASELINE configs, one line each (weak scaling like the headline) ------------------------------
    for other_name, steps in (("C1", 5), ("C2l", 3), ("C2t", 3), ("C3", 2), ("C4", 2), ("C5", 3), ("C5d", 3)):
        if other_name == name:
            continue
        per_gpu = {"C1": 4096, "C5": 1 << 18, "C5d": 1 << 18, "C4": 1 << 22}.get(other_name, N_PER_GPU)
        j = Job(args, other_name, per_gpu, world, rank, local)
        ms, total = timed(j, steps, 1, barrier, reduce_max_sum)
        if rank == 0:
            secondary[other_name] = line(j, ms, total)
            if j.is_traj:
                from clode_b200.flops import trajectory_bytes_per_point
                stored = int(j.sim.get_trajectory_counts().astype(np.int64).sum()) + j.n
                gbs = trajectory_bytes_per_point(j.w["model"], j.dims[2]) * stored / (ms * 1e-3) / 1e9
                peaks_file = os.path.join(REPO, "MEASURED_PEAKS.json")
                hbm = json.load(open(peaks_file)).get("hbm_gbs", 6650.0) if os.path.exists(peaks_file) else 6650.0
                secondary[other_name].update({"hbm_write_GBps_per_gpu": gbs, "frac_of_hbm_peak": gbs / hbm, "hbm_peak_GBps": hbm})
        j.close()

    # ---- through the user-facing Python API, numpy arrays in, ObserverOutput out ------------------------------------
    fe = frontend_leg(args, job, world, rank, local, dist)
    if rank == 0 and fe:
        top["e2e_frontend"] = fe
    if rank == 0 and world == 1:
        top["small_n"] = small_n_curve(local)
    return {"top": top, "secondary": secondary}


def frontend_leg(args, job, world, rank, local, dist):
    """FeatureSimulator.set_ensemble -> features() -> ObserverOutput with ordinary numpy arrays.  At N > 1 this is the
    in-process multi-GPU path: rank 0 alone drives all N GPUs (device_ids=[0..N-1]) while the other ranks wait on the
    rendezvous store (a CPU wait: an NCCL barrier would keep their GPUs busy)."""
    import torch

    store = None
    if dist:
        from torch.distributed import distributed_c10d
        store = distributed_c10d._get_default_store()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    result = None
    if rank == 0:
        try:
            import clode_b200 as clode
            from clode_b200 import build
            from clode_b200.models import rhs_path
            build.build_wrapper()
            w = workload(args.workload, job.n_total, np.arange(job.n_total))
            nv, npar, na, nw = job.dims
            var_names = ["x", "y", "z", "u", "v"][:nv]
            par_names = [f"p{k}" for k in range(npar)]
            obs = {"basic": clode.Observer.basic, "basicall": clode.Observer.basic_all_variables, "localmax": clode.Observer.local_max,
                   "thresh2": clode.Observer.threshold_2, "nhood2": clode.Observer.neighbourhood_2}[w["observer"]]
            stp = {"dopri5": clode.Stepper.dormand_prince, "bs23": clode.Stepper.bs23 if hasattr(clode.Stepper, "bs23") else clode.Stepper.bogacki_shampine,
                   "rk4": clode.Stepper.rk4, "seuler": clode.Stepper.stochastic_euler}[w["stepper"]]
            x0m = np.ascontiguousarray(w["x0"].reshape(nv, -1).T)      # the front end's (ensemble, nVar) matrices
            pm = np.ascontiguousarray(w["pars"].reshape(npar, -1).T)
            sp, opar = w["solver"], w["observer_params"]
            fs = clode.FeatureSimulator(src_file=rhs_path(w["model"]), variables=dict(zip(var_names, x0m[0])), parameters=dict(zip(par_names, pm[0])),
                                        aux=[f"a{k}" for k in range(na)], num_noise=nw, observer=obs, stepper=stp, single_precision=False,
                                        t_span=w["tspan"], dt=sp["dt"], dtmax=sp["dtmax"], abstol=sp["abstol"], reltol=sp["reltol"],
                                        max_steps=sp["max_steps"], observer_max_event_count=opar.get("max_event_count", 100),
                                        observer_x_up_thresh=opar.get("x_up_threshold", 0.3), observer_x_down_thresh=opar.get("x_down_threshold", 0.2),
                                        platform_id=0, device_ids=list(range(world)))
            split = {"set_ensemble_ms": 0.0, "features_and_results_ms": 0.0}

            def one():
                ta = time.perf_counter()
                fs.set_ensemble(variables=x0m, parameters=pm)
                tb = time.perf_counter()
                out = fs.features(initialize_observer=True, update_x0=False)
                tc = time.perf_counter()
                split["set_ensemble_ms"] += (tb - ta) * 1e3
                split["features_and_results_ms"] += (tc - tb) * 1e3
                return out
            for _ in range(2):
                out = one()
            split = {k: 0.0 for k in split}
            t0 = time.perf_counter()
            K = max(args.steps // 2, 3)
            for _ in range(K):
                out = one()
            sec = (time.perf_counter() - t0) / K
            split = {k: v / K for k, v in split.items()}
            steps_name = [f for f in out.get_feature_names() if "step" in f and "count" in f][0]
            total = int(np.asarray(out.F[steps_name]).sum())
            kernel_ms = fs._integrator.get_last_kernel_ms()
            result = {"value": total / sec, "unit": "instance-steps/s", "ms_per_step": sec * 1e3, "kernel_ms_per_step": kernel_ms,
                      "fraction_of_kernel_only": kernel_ms / (sec * 1e3), "n_gpus": world, "instances_total": job.n_total, "split": split,
                      "path": "clode_b200.FeatureSimulator(device_ids=[0..N-1]).set_ensemble(numpy) -> features() -> ObserverOutput (record array): "
                              "transposing upload through the page-locked staging ring, NVLink gather + GPU transpose, one D2H into a pooled "
                              "page-locked block that the record array views",
                      "h2d_bytes_per_step": 8 * job.n_total * (nv + npar), "d2h_bytes_per_step": 8 * job.n_total * fs._integrator.get_n_features()}
            del fs
        except Exception as e:  # the leg is a report, never a reason to lose the headline
            result = {"error": repr(e)[:400]}
        if store is not None:
            store.set("clode_frontend_leg_done", "1")
    elif store is not None:
        store.wait(["clode_frontend_leg_done"])
    return result


def small_n_curve(local):
    """the reference's only performance script (examples/dump_device_performance.py:36-105): Lorenz, rk4, 1000 steps,
    N = 2^0 .. 2^17, single precision; here per N the wall time of one transient() call and its kernel time"""
    from clode_b200 import _rt
    from clode_b200.models import MODELS, rhs_source

    nv, npar, na, nw = MODELS["lorenz63"]
    out = []
    for single in (True, False):
        prog = _rt.Program(rhs_source("lorenz63"), "rk4", nv, npar, na, nw, kernels=_rt.KERNEL_TRANSIENT, single_precision=single)
        t0 = time.perf_counter()
        sim = _rt.Sim(prog, device=local)
        build_s = time.perf_counter() - t0
        sim.set_solver_params(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=1000)
        sim.set_tspan(0.0, 10.0)
        rows = []
        for e in range(0, 18):
            n = 1 << e
            sim.set_problem(np.ones(nv * n), np.tile([28.0, 10.0, 8.0 / 3.0], (n, 1)).T.ravel())
            sim.transient()
            best_wall, best_ms = 1e9, 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                sim.transient()
                best_wall = min(best_wall, time.perf_counter() - t0)
                best_ms = min(best_ms, sim.last_kernel_ms())
            rows.append({"n": n, "call_us": best_wall * 1e6, "kernel_us": best_ms * 1e3, "instance_steps_per_s": n * 1000 / best_wall})
        out.append({"precision": "f32" if single else "f64", "program_build_or_cache_load_s": build_s, "steps_per_instance": 1000, "points": rows})
        sim.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--npts", type=int, default=0, help="instances per GPU (default 2^20)")
    ap.add_argument("--quick", type=int, default=0, help="1: headline + e2e only (A/B sweeps); 0: also tiers, parity, strong scaling, secondary configs, front end, small-N curve")
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
    ap.add_argument("--bit-exact", type=int, default=0, help="time the bit-exact tier as the headline instead of the production tier")
    ap.add_argument("--shuffle", type=int, default=0, help="randomly permute the parameter grid (heterogeneous warps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
