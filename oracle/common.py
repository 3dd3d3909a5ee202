"""TEST INFRASTRUCTURE — not part of the product.

Shared descriptions for the two CPU oracles (``oracle/_ref`` = the reference's own
kernel sources compiled as host C, ``oracle/restate`` = an independent C restatement):
model registry, build configuration, and ctypes mirrors of the two parameter structs
(`struct SolverParams` clode/cpp/clODE_struct_defs.cl:11-20, `struct ObserverParams`
clode/cpp/observers.cl:25-46).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS_DIR = os.path.join(REPO, "clode_b200", "models")
REFERENCE_ROOT = os.environ.get("CLODE_REFERENCE_ROOT", "/root/reference")

# model registry (dimensions + RHS files) lives with the product; the oracle only reads it
import sys as _sys

if REPO not in _sys.path:
    _sys.path.insert(0, REPO)
from clode_b200.models import MODELS  # noqa: E402  name -> (nVar, nPar, nAux, nWiener)

# the reference's own RHS files for the same systems (only present in the build
# container); used to check that clode_b200/models/*.cl are arithmetic-identical.
REFERENCE_RHS = {
    "lorenz63": "test/lorenz.cl",
    "vanderpol": "test/van_der_pol_oscillator.cl",
    "thompson_a1": "test/ornl_thompson_a1.cl",
    "lactotroph": "samples/lactotroph.cl",
    "lactotroph_noise": "samples/lactotroph_noise.cl",
    "chay_keizer": "examples/chay_keizer.cl",
}

# stepper name -> -D define (clode/cpp/steppers.cl:29-34)
STEPPER_DEFINES = {
    "euler": "EXPLICIT_EULER",
    "heun": "EXPLICIT_HEUN",
    "rk4": "EXPLICIT_RK4",
    "bs23": "EXPLICIT_BS23",
    "dopri5": "EXPLICIT_DOPRI5",
    "seuler": "STOCHASTIC_EULER",
}

# observer name -> -D define (clode/cpp/observers/*.clh, `oi.define=`)
OBSERVER_DEFINES = {
    "basic": "USE_OBSERVER_BASIC",
    "basicall": "USE_OBSERVER_BASIC_ALLVAR",
    "localmax": "USE_OBSERVER_LOCAL_MAX",
    "nhood1": "USE_OBSERVER_NEIGHBORHOOD_1",
    "nhood2": "USE_OBSERVER_NEIGHBORHOOD_2",
    "thresh2": "USE_OBSERVER_THRESHOLD_2",
}


def n_features(observer: str, n_var: int, n_aux: int, n_store: int) -> int:
    """Feature-vector length per observer (the `featureNames` lists built in
    clode/cpp/observers/observer_*.clh `getObserverInfo_*`)."""
    return {
        "basic": 6,
        "basicall": 5 * n_var + 3 * n_aux + 1,
        "localmax": 6 + 5 * n_var + 3 * n_aux + 4 * n_store + 2,
        "nhood1": 6 + 5 * n_var + 3 * n_aux + 5,
        "nhood2": 6 + 7 * n_var + 3 * n_aux + n_store + 5,
        "thresh2": 18 + 5 * n_var + 3 * n_aux + 2 * n_store + 5,
    }[observer]


@dataclass(frozen=True)
class Config:
    model: str
    stepper: str
    observer: str = "basic"
    n_store_events: int = 0
    single: bool = False
    math: str = "libm"  # "libm" (glibc) | "pm" (portable math, bit-exact tier)
    contract: str = "off"  # gcc -ffp-contract=
    rhs_path: str = ""  # override: absolute path of a getRHS file
    dims: tuple = ()  # override for models not in MODELS

    @property
    def shape(self):
        return tuple(self.dims) if self.dims else MODELS[self.model]

    @property
    def rhs_file(self):
        return self.rhs_path or os.path.join(MODELS_DIR, self.model + ".cl")

    @property
    def real(self):
        return np.float32 if self.single else np.float64

    @property
    def tag(self):
        return "_".join(
            [
                self.model,
                self.stepper,
                self.observer,
                f"ns{self.n_store_events}",
                "f32" if self.single else "f64",
                self.math,
                "c" + self.contract,
            ]
        )

    def defines(self):
        n_var, n_par, n_aux, n_wiener = self.shape
        d = [
            "-DCLODE_SINGLE_PRECISION" if self.single else "-DCLODE_DOUBLE_PRECISION",
            "-D" + STEPPER_DEFINES[self.stepper],
            f"-DN_PAR={n_par}",
            f"-DN_VAR={n_var}",
            f"-DN_AUX={n_aux}",
            f"-DN_WIENER={n_wiener}",
            "-D" + OBSERVER_DEFINES[self.observer],
            f"-DN_STORE_EVENTS={self.n_store_events}",
        ]
        return d


def _solver_params_type(real):
    class SolverParams(ctypes.Structure):
        _fields_ = [
            ("dt", real),
            ("dtmax", real),
            ("abstol", real),
            ("reltol", real),
            ("max_steps", ctypes.c_uint),
            ("max_store", ctypes.c_uint),
            ("nout", ctypes.c_uint),
        ]

    return SolverParams


def _observer_params_type(real):
    class ObserverParams(ctypes.Structure):
        _fields_ = [
            ("eVarIx", ctypes.c_uint),
            ("fVarIx", ctypes.c_uint),
            ("maxEventCount", ctypes.c_uint),
            ("maxEventTimestamps", ctypes.c_uint),
            ("minXamp", real),
            ("minIMI", real),
            ("nHoodRadius", real),
            ("xUpThresh", real),
            ("xDownThresh", real),
            ("dxUpThresh", real),
            ("dxDownThresh", real),
            ("eps_dx", real),
        ]

    return ObserverParams


SolverParamsD = _solver_params_type(ctypes.c_double)
SolverParamsF = _solver_params_type(ctypes.c_float)
ObserverParamsD = _observer_params_type(ctypes.c_double)
ObserverParamsF = _observer_params_type(ctypes.c_float)


@dataclass
class Solver:
    """python-side mirror of SolverParams with the pybind defaults
    (clode/cpp/CLODEpython.cpp:229-235)."""

    dt: float = 0.1
    dtmax: float = 0.5
    abstol: float = 1e-6
    reltol: float = 1e-3
    max_steps: int = 1000000
    max_store: int = 1000000
    nout: int = 1

    def c(self, single=False):
        T = SolverParamsF if single else SolverParamsD
        return T(self.dt, self.dtmax, self.abstol, self.reltol, self.max_steps, self.max_store, self.nout)


@dataclass
class Observer:
    """python-side mirror of ObserverParams with the pybind defaults
    (clode/cpp/CLODEpython.cpp:302-313)."""

    e_var_ix: int = 0
    f_var_ix: int = 0
    max_event_count: int = 100
    max_event_timestamps: int = 0
    min_amp: float = 0.0
    min_imi: float = 0.0
    nhood_radius: float = 0.05
    x_up_threshold: float = 0.2
    x_down_threshold: float = 0.2
    dx_up_threshold: float = 0.0
    dx_down_threshold: float = 0.0
    eps_dx: float = 0.0

    def c(self, single=False):
        T = ObserverParamsF if single else ObserverParamsD
        return T(
            self.e_var_ix,
            self.f_var_ix,
            self.max_event_count,
            self.max_event_timestamps,
            self.min_amp,
            self.min_imi,
            self.nhood_radius,
            self.x_up_threshold,
            self.x_down_threshold,
            self.dx_up_threshold,
            self.dx_down_threshold,
            self.eps_dx,
        )


def seed_states(seed: int, n_pts: int, offset: int = 0, n_global: int | None = None) -> np.ndarray:
    """RNG seeding rule of `CLODE::seedRNG(cl_int)` (clode/cpp/CLODE.cpp:447-453):
    RNGstate[k] = seed + k for k in [0, 2*nPts); instance i owns words i and nPts+i.
    `offset`/`n_global` give the slice a shard [offset, offset+n_pts) of a larger
    ensemble must use to reproduce the unsharded stream."""
    n_global = n_pts if n_global is None else n_global
    i = np.arange(offset, offset + n_pts, dtype=np.int64)
    # `mySeed + i` is evaluated in cl_int (32-bit, wraps) and then widened to cl_ulong (sign extension)
    wrap = lambda k: (np.int64(seed) + k).astype(np.int32).astype(np.int64).astype(np.uint64)
    return np.concatenate([wrap(i), wrap(np.int64(n_global) + i)])
