"""TEST INFRASTRUCTURE — not part of the product.

oracle/_ref: the reference's OWN kernel sources, compiled as host C.

`build(cfg)` compiles `oracle/ref_driver.c`, which #includes the reference's
transient.cl / initializeObserver.cl / features.cl / trajectory.cl from the reference
tree, into `oracle/_ref/ref_<tag>.so`.  The .so files are git-ignored but travel to the
GPU box with the repository snapshot; the reference tree itself is only present in the
build container, so `load(cfg)` never needs it.

Three documented source patches are applied to a throw-away copy of clode/cpp (under
oracle/_ref/, deleted after the compile) because the unpatched code is memory-unsafe
as host C:
  D1  observer_local_maximum.clh:278  `tMinList[eventcount-1]` written with
      eventcount == 0 (index 0xFFFFFFFF)            -> guard with eventcount > 0
  D2  observer_threshold_2.clh:352,377 and observer_neighborhood_1.clh:269:
      `thisXbuffer[N_VAR]` filled with 3 elements   -> size 3
Nothing else differs from the reference sources.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import tempfile

import numpy as np

from .common import REFERENCE_ROOT, REPO, Config, Observer, Solver, n_features

REF_DIR = os.path.join(REPO, "oracle", "_ref")
HERE = os.path.dirname(os.path.abspath(__file__))
PM_INCLUDE = os.path.join(REPO, "clode_b200", "csrc", "device")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "clode", "cpp", "transient.cl"))


def so_path(cfg: Config) -> str:
    return os.path.join(REF_DIR, f"ref_{cfg.tag}.so")


_PATCHES = [
    # (file, old, new, minimum occurrences; commented-out copies are patched too)
    (
        "observers/observer_local_maximum.clh",
        "        if (od->eventcount <= N_STORE_EVENTS) {\n            od->tMinList[od->eventcount-1]",
        "        if (od->eventcount > 0 && od->eventcount <= N_STORE_EVENTS) {\n            od->tMinList[od->eventcount-1]",
        1,
    ),
    ("observers/observer_threshold_2.clh", "realtype thisXbuffer[N_VAR];", "realtype thisXbuffer[3];", 2),
    ("observers/observer_neighborhood_1.clh", "realtype thisXbuffer[N_VAR];", "realtype thisXbuffer[3];", 1),
]


def _patched_tree(dst: str) -> str:
    src = os.path.join(REFERENCE_ROOT, "clode", "cpp")
    out = os.path.join(dst, "cpp")
    shutil.copytree(src, out, ignore=shutil.ignore_patterns("OpenCL", "*.cpp", "*.hpp", "*.pyi", "BUILD", "logging"))
    for rel, old, new, count in _PATCHES:
        p = os.path.join(out, rel)
        text = open(p).read()
        if text.count(old) < count:
            raise RuntimeError(f"reference patch site moved: {rel} ({text.count(old)} < {count})")
        open(p, "w").write(text.replace(old, new))
    return out


def build(cfg: Config, force: bool = False, opt: str = "-O2") -> str:
    """Compile the reference kernels for one configuration. Requires the reference tree."""
    out = so_path(cfg)
    if os.path.exists(out) and not force:
        return out
    if not reference_available():
        raise FileNotFoundError(f"{out} is not built and the reference tree is absent")
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix=".src_", dir=REF_DIR)
    try:
        inc = _patched_tree(tmp)
        cmd = [
            "gcc", "-x", "c", "-std=gnu11", opt, "-march=x86-64-v3", f"-ffp-contract={cfg.contract}",
            "-fno-math-errno", "-fopenmp", "-shared", "-fPIC", "-w",
            f"-I{inc}", f"-I{HERE}", f"-I{PM_INCLUDE}",
            f'-DREF_RHS_FILE="{cfg.rhs_file}"',
            *cfg.defines(),
        ]
        if cfg.math == "pm":
            cmd.append("-DREF_PM_MATH")
        cmd += [os.path.join(HERE, "ref_driver.c"), "-o", out, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + " ".join(cmd) + "\n" + r.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RefLib:
    """ctypes front for one compiled configuration; arrays use the reference's flat
    variable-major layout (x0[j*nPts+i], F[k*nPts+i], x[s*nPts*nVar + j*nPts + i])."""

    def __init__(self, cfg: Config):
        path = so_path(cfg)
        if not os.path.exists(path):
            build(cfg)
        self.cfg = cfg
        self.lib = ctypes.CDLL(path)
        info = (ctypes.c_long * 8)()
        self.lib.ref_info(info)
        self.real_size, self.odata_size = info[0], info[1]
        self.n_var, self.n_par, self.n_aux, self.n_wiener = info[4], info[5], info[6], info[7]
        assert (self.n_var, self.n_par, self.n_aux, self.n_wiener) == tuple(cfg.shape)
        self.real = cfg.real
        self.n_feat = n_features(cfg.observer, self.n_var, self.n_aux, cfg.n_store_events)
        self.odata = None

    # -- helpers -------------------------------------------------------------
    def _prep(self, x0, pars, dt, rng, n_pts):
        x0 = np.ascontiguousarray(x0, dtype=self.real).copy()
        pars = np.ascontiguousarray(pars, dtype=self.real).copy()
        assert x0.size == self.n_var * n_pts and pars.size == self.n_par * n_pts
        dt = np.ascontiguousarray(dt, dtype=self.real).copy()
        rng = np.ascontiguousarray(rng, dtype=np.uint64).copy()
        assert dt.size == n_pts and rng.size == 2 * n_pts
        return x0, pars, dt, rng

    def transient(self, tspan, x0, pars, sp: Solver, dt, rng, nthreads=1):
        n = len(dt)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        xf = np.zeros(self.n_var * n, self.real)
        tf = np.zeros(n, self.real)
        spc = sp.c(self.cfg.single)
        self.lib.ref_transient(n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars), ctypes.byref(spc),
                               _ptr(xf), _ptr(rng), _ptr(dt), _ptr(tf))
        return dict(xf=xf, tf=tf, dt=dt, rng=rng)

    def initialize_observer(self, tspan, x0, pars, sp: Solver, op: Observer, dt, rng, nthreads=1):
        n = len(dt)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        self.odata = np.zeros(self.odata_size * n + 64, np.uint8)
        spc, opc = sp.c(self.cfg.single), op.c(self.cfg.single)
        self.lib.ref_initialize_observer(n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars), ctypes.byref(spc),
                                         _ptr(rng), _ptr(dt), _ptr(self.odata), ctypes.byref(opc))

    def features(self, tspan, x0, pars, sp: Solver, op: Observer, dt, rng, initialize=True, nthreads=1):
        """`CLODEfeatures::features()` (clode/cpp/CLODEfeatures.cpp:222-258): runs
        initializeObserver first unless the observer data is being continued."""
        n = len(dt)
        if initialize or self.odata is None:
            self.initialize_observer(tspan, x0, pars, sp, op, dt, rng, nthreads)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        xf = np.zeros(self.n_var * n, self.real)
        tf = np.zeros(n, self.real)
        F = np.zeros(self.n_feat * n, self.real)
        spc, opc = sp.c(self.cfg.single), op.c(self.cfg.single)
        self.lib.ref_features(n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars), ctypes.byref(spc), _ptr(xf),
                              _ptr(rng), _ptr(dt), _ptr(tf), _ptr(self.odata), ctypes.byref(opc), _ptr(F))
        return dict(F=F, xf=xf, tf=tf, dt=dt, rng=rng)

    def trajectory(self, tspan, x0, pars, sp: Solver, dt, rng, nthreads=1):
        n = len(dt)
        x0, pars, dt, rng = self._prep(x0, pars, dt, rng, n)
        ts = np.asarray(tspan, dtype=self.real)
        xf = np.zeros(self.n_var * n, self.real)
        tf = np.zeros(n, self.real)
        rows = sp.max_store + 1  # the kernel can write row index max_store
        t = np.zeros(rows * n, self.real)
        x = np.zeros(rows * n * self.n_var, self.real)
        dx = np.zeros(rows * n * self.n_var, self.real)
        aux = np.zeros(max(1, rows * n * self.n_aux), self.real)
        nst = np.zeros(n, np.int32)
        spc = sp.c(self.cfg.single)
        self.lib.ref_trajectory(n, nthreads, _ptr(ts), _ptr(x0), _ptr(pars), ctypes.byref(spc), _ptr(xf),
                                _ptr(rng), _ptr(dt), _ptr(tf), _ptr(t), _ptr(x), _ptr(dx), _ptr(aux), _ptr(nst))
        return dict(t=t, x=x, dx=dx, aux=aux, n_stored=nst, xf=xf, tf=tf, dt=dt, rng=rng, rows=rows)
