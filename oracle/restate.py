"""TEST INFRASTRUCTURE — not part of the product.

ctypes front for the C restatement oracle (oracle/restate/clode_oracle.c).  One shared
object per (right-hand side, precision, math flavour, contraction); it is compiled on
demand with gcc, which exists both in the build container and on the GPU box, so this
oracle needs neither the reference tree nor a prebuilt artefact.  The call signatures
match oracle.ref.RefLib so tests can swap the two.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np

from .common import REPO, Config, Observer, Solver, n_features

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(REPO, "oracle", "_build")
PM_INCLUDE = os.path.join(REPO, "clode_b200", "csrc", "device")
SRC = os.path.join(HERE, "restate", "clode_oracle.c")

STEPPERS = ["euler", "heun", "rk4", "bs23", "dopri5", "seuler"]
OBSERVERS = ["basic", "basicall", "localmax", "nhood1", "nhood2", "thresh2"]


class _Problem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n_var", "n_par", "n_aux", "n_wiener", "stepper", "observer", "n_store")]


def _so_path(cfg: Config) -> str:
    h = hashlib.sha1()
    h.update(open(SRC, "rb").read())
    h.update(open(cfg.rhs_file, "rb").read())
    h.update(open(os.path.join(PM_INCLUDE, "pm_math.h"), "rb").read())
    name = os.path.splitext(os.path.basename(cfg.rhs_file))[0]
    tag = f"{name}_{'f32' if cfg.single else 'f64'}_{cfg.math}_c{cfg.contract}_{h.hexdigest()[:10]}"
    return os.path.join(BUILD_DIR, f"oracle_{tag}.so")


def build(cfg: Config, opt: str = "-O2") -> str:
    out = _so_path(cfg)
    if os.path.exists(out):
        return out
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = [
        "gcc", "-std=gnu11", opt, "-march=x86-64-v3", f"-ffp-contract={cfg.contract}", "-fno-math-errno",
        "-fopenmp", "-shared", "-fPIC", "-w", f"-I{PM_INCLUDE}", f'-DOR_RHS_FILE="{cfg.rhs_file}"',
    ]
    if cfg.single:
        cmd.append("-DOR_SINGLE")
    if cfg.math == "pm":
        cmd.append("-DOR_PM_MATH")
    tmp = out + f".tmp{os.getpid()}"
    cmd += [SRC, "-o", tmp, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + " ".join(cmd) + "\n" + r.stderr)
    os.replace(tmp, out)
    return out


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleLib:
    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.lib = ctypes.CDLL(build(cfg))
        self.lib.or_obs_size.restype = ctypes.c_long
        self.n_var, self.n_par, self.n_aux, self.n_wiener = cfg.shape
        self.real = cfg.real
        self.n_feat = n_features(cfg.observer, self.n_var, self.n_aux, cfg.n_store_events)
        self.pb = _Problem(self.n_var, self.n_par, self.n_aux, self.n_wiener, STEPPERS.index(cfg.stepper),
                           OBSERVERS.index(cfg.observer), cfg.n_store_events)
        self.obs_size = self.lib.or_obs_size()
        self.odata = None

    def _prep(self, x0, pars, dt, rng, n):
        x0 = np.ascontiguousarray(x0, dtype=self.real).copy()
        pars = np.ascontiguousarray(pars, dtype=self.real).copy()
        assert x0.size == self.n_var * n and pars.size == self.n_par * n
        dt = np.ascontiguousarray(dt, dtype=self.real).copy()
        rng = np.ascontiguousarray(rng, dtype=np.uint64).copy()
        assert dt.size == n and rng.size == 2 * n
        return x0, pars, dt, rng

    def transient(self, tspan, x0, pars, sp: Solver, dt, rng, nthreads=1):
        n = len(dt)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        xf = np.zeros(self.n_var * n, self.real)
        tf = np.zeros(n, self.real)
        spc = sp.c(self.cfg.single)
        self.lib.or_transient(ctypes.byref(self.pb), n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars),
                              ctypes.byref(spc), _ptr(xf), _ptr(rng), _ptr(dt), _ptr(tf))
        return dict(xf=xf, tf=tf, dt=dt, rng=rng)

    def initialize_observer(self, tspan, x0, pars, sp: Solver, op: Observer, dt, rng, nthreads=1):
        n = len(dt)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        self.odata = np.zeros(self.obs_size * n, np.uint8)
        spc, opc = sp.c(self.cfg.single), op.c(self.cfg.single)
        self.lib.or_initialize_observer(ctypes.byref(self.pb), n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars),
                                        ctypes.byref(spc), _ptr(rng), _ptr(dt), _ptr(self.odata),
                                        ctypes.byref(opc))

    def features(self, tspan, x0, pars, sp: Solver, op: Observer, dt, rng, initialize=True, nthreads=1):
        n = len(dt)
        if initialize or self.odata is None:
            self.initialize_observer(tspan, x0, pars, sp, op, dt, rng, nthreads)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        xf = np.zeros(self.n_var * n, self.real)
        tf = np.zeros(n, self.real)
        F = np.zeros(self.n_feat * n, self.real)
        spc, opc = sp.c(self.cfg.single), op.c(self.cfg.single)
        self.lib.or_features(ctypes.byref(self.pb), n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars),
                             ctypes.byref(spc), _ptr(xf), _ptr(rng), _ptr(dt), _ptr(tf), _ptr(self.odata),
                             ctypes.byref(opc), _ptr(F))
        return dict(F=F, xf=xf, tf=tf, dt=dt, rng=rng)

    def trajectory(self, tspan, x0, pars, sp: Solver, dt, rng, nthreads=1):
        n = len(dt)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        xf = np.zeros(self.n_var * n, self.real)
        tf = np.zeros(n, self.real)
        rows = sp.max_store + 1
        t = np.zeros(rows * n, self.real)
        x = np.zeros(rows * n * self.n_var, self.real)
        dx = np.zeros(rows * n * self.n_var, self.real)
        aux = np.zeros(max(1, rows * n * self.n_aux), self.real)
        nst = np.zeros(n, np.int32)
        spc = sp.c(self.cfg.single)
        self.lib.or_trajectory(ctypes.byref(self.pb), n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars),
                               ctypes.byref(spc), _ptr(xf), _ptr(rng), _ptr(dt), _ptr(tf), _ptr(t), _ptr(x),
                               _ptr(dx), _ptr(aux), _ptr(nst))
        return dict(t=t, x=x, dx=dx, aux=aux, n_stored=nst, xf=xf, tf=tf, dt=dt, rng=rng, rows=rows)
