"""TEST INFRASTRUCTURE: run the product's device code on the CPU (see cuda_emu.h).

`EmuLib(cfg)` compiles tests/emu/emu_driver.cpp — i.e. clode_b200/csrc/device/*.cuh plus a
model RHS — with g++ for one configuration and exposes the same call signatures as
oracle.restate.OracleLib, so CPU tests can diff the device code against the oracle.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
DEVICE = os.path.join(REPO, "clode_b200", "csrc", "device")
BUILD = os.path.join(REPO, "oracle", "_build")

import sys
if REPO not in sys.path:
    sys.path.insert(0, REPO)
from oracle.common import OBSERVER_DEFINES, STEPPER_DEFINES, Config, Observer, Solver, n_features  # noqa: E402


class KernelArgs(ctypes.Structure):
    """mirror of struct KernelArgs in csrc/device/kernels.cuh"""
    _fields_ = [
        ("t0", ctypes.c_double), ("t1", ctypes.c_double),
        ("sp_dt", ctypes.c_double), ("sp_dtmax", ctypes.c_double), ("sp_abstol", ctypes.c_double),
        ("sp_reltol", ctypes.c_double),
        ("sp_max_steps", ctypes.c_uint), ("sp_max_store", ctypes.c_uint), ("sp_nout", ctypes.c_uint),
        ("op_max_event_count", ctypes.c_uint),
        ("op_min_x_amp", ctypes.c_double), ("op_min_imi", ctypes.c_double), ("op_nhood_radius", ctypes.c_double),
        ("op_x_up", ctypes.c_double), ("op_x_down", ctypes.c_double), ("op_dx_up", ctypes.c_double),
        ("op_dx_down", ctypes.c_double), ("op_eps_dx", ctypes.c_double),
        ("n", ctypes.c_ulonglong),
    ] + [(k, ctypes.c_void_p) for k in ("x0", "pars", "xf", "rng", "dt", "tf", "steps", "od_real", "od_uint", "F",
                                        "tr_t", "tr_x", "tr_dx", "tr_aux", "n_stored", "queue",
                                        "rs_real", "rs_uint", "chunk_flags")] + [
        ("row_begin", ctypes.c_uint), ("row_end", ctypes.c_uint), ("resume", ctypes.c_uint), ("block_order", ctypes.c_uint),
        ("cost_in", ctypes.c_void_p), ("cost_out", ctypes.c_void_p),
        ("perm", ctypes.c_void_p), ("n_slots", ctypes.c_ulonglong), ("attempt_budget", ctypes.c_uint), ("sched_resume", ctypes.c_uint),
        ("park_real", ctypes.c_void_p), ("park_uint", ctypes.c_void_p), ("sched_state", ctypes.c_void_p)]


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class EmuLib:
    def __init__(self, cfg: Config, f_var_ix=0, e_var_ix=0, extra_defs=()):
        self.cfg = cfg
        self.n_var, self.n_par, self.n_aux, self.n_wiener = cfg.shape
        self.real = cfg.real
        defs = [
            "-DCLODE_SINGLE_PRECISION" if cfg.single else "-DCLODE_DOUBLE_PRECISION",
            "-D" + STEPPER_DEFINES[cfg.stepper], "-D" + OBSERVER_DEFINES[cfg.observer],
            f"-DN_VAR={self.n_var}", f"-DN_PAR={self.n_par}", f"-DN_AUX={self.n_aux}", f"-DN_WIENER={self.n_wiener}",
            f"-DN_STORE_EVENTS={cfg.n_store_events}", f"-DF_VAR_IX={f_var_ix}", f"-DE_VAR_IX={e_var_ix}",
            "-DCLODE_WITH_FEATURES", "-DCLODE_WITH_TRAJECTORY", f'-DEMU_RHS_FILE="{cfg.rhs_file}"',
        ]
        defs += list(extra_defs)  # e.g. the shared-memory placement of the observer extents (observers.cuh)
        if cfg.math == "pm":
            defs.append("-DCLODE_BITEXACT")
        else:
            defs.append("-DCLODE_REFERENCE_MATH")  # libm flavour: no performance substitutions (controller root)
        h = hashlib.sha1(" ".join(defs).encode())
        for f in sorted(os.listdir(DEVICE)) + ["../../../tests/emu/cuda_emu.h", "../../../tests/emu/emu_driver.cpp"]:
            h.update(open(os.path.join(DEVICE, f), "rb").read())
        h.update(open(cfg.rhs_file, "rb").read())
        os.makedirs(BUILD, exist_ok=True)
        so = os.path.join(BUILD, f"emu_{cfg.tag}_{h.hexdigest()[:10]}.so")  # the digest covers extra_defs
        if not os.path.exists(so):
            cmd = ["g++", "-std=c++17", "-O2", "-march=x86-64-v3", f"-ffp-contract={cfg.contract}", "-fno-math-errno",
                   "-fPIC", "-shared", "-w", f"-I{HERE}", f"-I{DEVICE}", *defs,
                   os.path.join(HERE, "emu_driver.cpp"), "-o", so + f".tmp{os.getpid()}"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("emu build failed:\n" + " ".join(cmd) + "\n" + r.stderr)
            os.replace(so + f".tmp{os.getpid()}", so)  # atomic: concurrent test workers may build the same library
        self.lib = ctypes.CDLL(so)
        assert self.lib.emu_args_size() == ctypes.sizeof(KernelArgs)
        lay = (ctypes.c_int * 3)()
        self.lib.emu_observer_layout(lay)
        self.od_nreal, self.od_nuint, self.two_pass = lay[0], lay[1], lay[2]
        self.n_feat = n_features(cfg.observer, self.n_var, self.n_aux, cfg.n_store_events)
        self.od_real = self.od_uint = None

    def _args(self, tspan, sp: Solver, op: Observer, n, bufs):
        a = KernelArgs()
        a.t0, a.t1 = tspan
        a.sp_dt, a.sp_dtmax, a.sp_abstol, a.sp_reltol = sp.dt, sp.dtmax, sp.abstol, sp.reltol
        a.sp_max_steps, a.sp_max_store, a.sp_nout = sp.max_steps, sp.max_store, sp.nout
        if op is not None:
            a.op_max_event_count = op.max_event_count
            a.op_min_x_amp, a.op_min_imi, a.op_nhood_radius = op.min_amp, op.min_imi, op.nhood_radius
            a.op_x_up, a.op_x_down, a.op_dx_up, a.op_dx_down = (op.x_up_threshold, op.x_down_threshold,
                                                                 op.dx_up_threshold, op.dx_down_threshold)
            a.op_eps_dx = op.eps_dx
        a.n = n
        a.row_begin, a.row_end, a.resume = 0, 0xFFFFFFFF, 0
        a.attempt_budget = 0xFFFFFFFF
        a.block_order = getattr(self, "block_order", 0)
        for k, v in bufs.items():
            setattr(a, k, _ptr(v) if v is not None else None)
        return a

    def _prep(self, x0, pars, dt, rng, n):
        x0 = np.ascontiguousarray(x0, dtype=self.real).copy()
        pars = np.ascontiguousarray(pars, dtype=self.real).copy()
        dt = np.ascontiguousarray(dt, dtype=self.real).copy()
        rng = np.ascontiguousarray(rng, dtype=np.uint64).copy()
        base = dict(x0=x0, pars=pars, dt=dt, rng=rng, xf=np.zeros(self.n_var * n, self.real),
                    tf=np.zeros(n, self.real), steps=np.zeros(n, np.uint32))
        return base

    def transient(self, tspan, x0, pars, sp, dt, rng, nthreads=1):
        n = len(dt)
        b = self._prep(x0, pars, dt, rng, n)
        a = self._args(tspan, sp, None, n, b)
        self.lib.emu_transient(ctypes.byref(a))
        return dict(xf=b["xf"], tf=b["tf"], dt=b["dt"], rng=b["rng"], steps=b["steps"])

    def initialize_observer(self, tspan, x0, pars, sp, op, dt, rng, nthreads=1):
        n = len(dt)
        b = self._prep(x0, pars, dt, rng, n)
        self.od_real = np.zeros(max(1, self.od_nreal * n), self.real)
        self.od_uint = np.zeros(max(1, self.od_nuint * n), np.uint32)
        a = self._args(tspan, sp, op, n, dict(b, od_real=self.od_real, od_uint=self.od_uint))
        self.lib.emu_initialize_observer(ctypes.byref(a))

    def features(self, tspan, x0, pars, sp, op, dt, rng, initialize=True, nthreads=1):
        n = len(dt)
        if initialize or self.od_real is None:
            self.initialize_observer(tspan, x0, pars, sp, op, dt, rng)
        b = self._prep(x0, pars, dt, rng, n)
        F = np.zeros(self.n_feat * n, self.real)
        a = self._args(tspan, sp, op, n, dict(b, od_real=self.od_real, od_uint=self.od_uint, F=F))
        self.lib.emu_features(ctypes.byref(a))
        return dict(F=F, xf=b["xf"], tf=b["tf"], dt=b["dt"], rng=b["rng"], steps=b["steps"])

    # ---- the chunked (scheduled) time loops, driven the way clode_rt.cpp's run_loop drives them -----------------
    def _chunked(self, kernel, a, n, budgets, seed):
        """pilot + rounds: after every launch the unfinished instances are re-ordered (here: randomly — the result must
        not depend on the order) and continued with the next attempt budget; the last round runs to completion"""
        nv, na = self.n_var, max(self.n_aux, 1)
        park_real = np.zeros((2 * nv + na + 3) * n, self.real)
        park_uint = np.zeros(2 * n, np.uint32)
        a.park_real, a.park_uint = _ptr(park_real), _ptr(park_uint)
        rng = np.random.default_rng(seed)
        a.attempt_budget, a.sched_resume, a.perm, a.n_slots = budgets[0], 0, None, 0
        kernel(ctypes.byref(a))
        launches = 1
        for k, budget in enumerate(list(budgets[1:]) + [0xFFFFFFFF]):
            live = np.flatnonzero((park_uint[n:] & 2) == 0).astype(np.uint32)
            if live.size == 0:
                break
            perm = rng.permutation(live).astype(np.uint32)
            a.perm, a.n_slots, a.attempt_budget, a.sched_resume = _ptr(perm), perm.size, budget, 1
            kernel(ctypes.byref(a))
            launches += 1
        assert np.all(park_uint[n:] & 2), "instances left unfinished"
        return launches, park_uint[:n].copy()

    def transient_chunked(self, tspan, x0, pars, sp, dt, rng, budgets=(7, 50, 300), seed=0):
        n = len(dt)
        b = self._prep(x0, pars, dt, rng, n)
        a = self._args(tspan, sp, None, n, b)
        launches, _ = self._chunked(self.lib.emu_transient, a, n, budgets, seed)
        return dict(xf=b["xf"], tf=b["tf"], dt=b["dt"], rng=b["rng"], steps=b["steps"], launches=launches)

    def features_chunked(self, tspan, x0, pars, sp, op, dt, rng, budgets=(7, 50, 300), seed=0):
        n = len(dt)
        b = self._prep(x0, pars, dt, rng, n)
        self.od_real = np.zeros(max(1, self.od_nreal * n), self.real)
        self.od_uint = np.zeros(max(1, self.od_nuint * n), np.uint32)
        a = self._args(tspan, sp, op, n, dict(b, od_real=self.od_real, od_uint=self.od_uint))
        launches = 1
        if self.two_pass:
            launches, warm_steps = self._chunked(self.lib.emu_initialize_observer, a, n, budgets, seed)
        else:
            self.lib.emu_initialize_observer(ctypes.byref(a))
        b = self._prep(x0, pars, dt, rng, n)
        F = np.zeros(self.n_feat * n, self.real)
        a = self._args(tspan, sp, op, n, dict(b, od_real=self.od_real, od_uint=self.od_uint, F=F))
        k, _ = self._chunked(self.lib.emu_features, a, n, budgets, seed + 1)
        return dict(F=F, xf=b["xf"], tf=b["tf"], dt=b["dt"], rng=b["rng"], steps=b["steps"], launches=launches + k)

    def trajectory(self, tspan, x0, pars, sp, dt, rng, nthreads=1):
        n = len(dt)
        b = self._prep(x0, pars, dt, rng, n)
        rows = sp.max_store + 1
        t = np.zeros(rows * n, self.real)
        x = np.zeros(rows * n * self.n_var, self.real)
        dx = np.zeros(rows * n * self.n_var, self.real)
        aux = np.zeros(max(1, rows * n * self.n_aux), self.real)
        nst = np.zeros(n, np.int32)
        a = self._args(tspan, sp, None, n, dict(b, tr_t=t, tr_x=x, tr_dx=dx, tr_aux=aux, n_stored=nst))
        self.lib.emu_trajectory(ctypes.byref(a))
        return dict(t=t, x=x, dx=dx, aux=aux, n_stored=nst, xf=b["xf"], tf=b["tf"], dt=b["dt"], rng=b["rng"],
                    rows=rows, steps=b["steps"])

    def trajectory_stream(self, tspan, x0, pars, sp, dt, rng, chunk_rows):
        """the chunked trajectory exactly as clode_sim_trajectory_stream drives it (clode_rt.cpp): launches of
        `chunk_rows` stored points into a chunk-sized buffer, state carried in xf/tf/dt/rng + rs_real/rs_uint"""
        n = len(dt)
        b = self._prep(x0, pars, dt, rng, n)
        rows, total = sp.max_store, sp.max_store + 1
        R = min(chunk_rows, total)
        nv, na = self.n_var, self.n_aux
        out = dict(t=np.zeros(rows * n, self.real), x=np.zeros(rows * n * nv, self.real),
                   dx=np.zeros(rows * n * nv, self.real), aux=np.zeros(max(1, rows * n * na), self.real))
        nst = np.zeros(n, np.int32)
        rs_real = np.zeros((1 + self.n_wiener) * n, self.real)
        rs_uint = np.zeros(3 * n, np.uint32)
        launches = 0
        for k in range((total + R - 1) // R):
            row_begin, row_end = k * R, min(k * R + R, total)
            c = dict(t=np.zeros(R * n, self.real), x=np.zeros(R * n * nv, self.real), dx=np.zeros(R * n * nv, self.real),
                     aux=np.zeros(max(1, R * n * na), self.real))
            flags = np.zeros(2, np.uint32)
            a = self._args(tspan, sp, None, n, dict(b, tr_t=c["t"], tr_x=c["x"], tr_dx=c["dx"], tr_aux=c["aux"], n_stored=nst,
                                                    rs_real=rs_real, rs_uint=rs_uint, chunk_flags=flags))
            a.row_begin, a.row_end, a.resume = row_begin, row_end, int(k > 0)
            self.lib.emu_trajectory(ctypes.byref(a))
            launches += 1
            stop = min(int(flags[1]) + 1, row_end, rows)
            for key, width in (("t", 1), ("x", nv), ("dx", nv), ("aux", na)):
                if stop > row_begin and width:
                    out[key][row_begin * width * n:stop * width * n] = c[key][:(stop - row_begin) * width * n]
            if not flags[0]:
                break
        return dict(out, n_stored=nst, xf=b["xf"], tf=b["tf"], dt=b["dt"], rng=b["rng"], rows=rows, steps=b["steps"],
                    launches=launches)
