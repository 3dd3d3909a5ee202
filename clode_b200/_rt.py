"""ctypes binding of the C ABI in include/clode_rt.h (libclode_rt.so).

This is the zero-copy numpy path into the runtime: arrays go straight from numpy buffers
to the device, without the list <-> std::vector conversion of the pybind layer
(SURVEY.md §8f-1).  The product never falls back to a CPU implementation: if the library
or the CUDA driver is missing, calls raise `RuntimeError`.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclode_rt.so")

KERNEL_TRANSIENT, KERNEL_FEATURES, KERNEL_TRAJECTORY = 1, 2, 4
(BUF_X0, BUF_PARS, BUF_XF, BUF_DT, BUF_TF, BUF_F, BUF_T, BUF_X, BUF_DX, BUF_AUX, BUF_RNG, BUF_STEPS,
 BUF_NSTORED) = range(13)


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"clode_rt error {code}: {msg}")
        self.code = code


class ProgramDesc(ctypes.Structure):
    _fields_ = [
        ("rhs_source", ctypes.c_char_p), ("stepper", ctypes.c_char_p), ("observer", ctypes.c_char_p),
        ("single_precision", ctypes.c_int), ("n_var", ctypes.c_int), ("n_par", ctypes.c_int),
        ("n_aux", ctypes.c_int), ("n_wiener", ctypes.c_int), ("f_var_ix", ctypes.c_int),
        ("e_var_ix", ctypes.c_int), ("n_store_events", ctypes.c_int), ("kernels", ctypes.c_int),
        ("bit_exact", ctypes.c_int), ("work_queue", ctypes.c_int), ("block_size", ctypes.c_int),
        ("min_blocks_per_sm", ctypes.c_int), ("staged_trajectory", ctypes.c_int),
        ("observer_in_shared", ctypes.c_int), ("ieee_constant_division", ctypes.c_int),
        ("library_exp", ctypes.c_int),
    ]


class SolverParamsC(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_double), ("dtmax", ctypes.c_double), ("abstol", ctypes.c_double),
                ("reltol", ctypes.c_double), ("max_steps", ctypes.c_uint), ("max_store", ctypes.c_uint),
                ("nout", ctypes.c_uint)]


class ObserverParamsC(ctypes.Structure):
    _fields_ = [("e_var_ix", ctypes.c_uint), ("f_var_ix", ctypes.c_uint), ("max_event_count", ctypes.c_uint),
                ("max_event_timestamps", ctypes.c_uint), ("min_x_amp", ctypes.c_double),
                ("min_imi", ctypes.c_double), ("nhood_radius", ctypes.c_double),
                ("x_up_thresh", ctypes.c_double), ("x_down_thresh", ctypes.c_double),
                ("dx_up_thresh", ctypes.c_double), ("dx_down_thresh", ctypes.c_double),
                ("eps_dx", ctypes.c_double)]


class DeviceInfoC(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 256), ("cc_major", ctypes.c_int), ("cc_minor", ctypes.c_int),
                ("multiprocessors", ctypes.c_int), ("clock_mhz", ctypes.c_int),
                ("max_threads_per_block", ctypes.c_int), ("total_memory", ctypes.c_uint64),
                ("max_alloc", ctypes.c_uint64), ("driver_version", ctypes.c_int)]


class KernelInfoC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("registers", "local_bytes", "shared_bytes", "const_bytes",
                                            "max_threads", "block_size", "blocks_per_sm", "grid_size")]


_lib = None


def lib():
    """Load libclode_rt.so (building it is `python -m clode_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m clode_b200.build`")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.clode_last_error.restype = ctypes.c_char_p
        _lib.clode_version.restype = ctypes.c_char_p
        _lib.clode_free.argtypes = [ctypes.c_void_p]
        _lib.clode_free.restype = None
    return _lib


def _check(rc):
    if rc != 0:
        raise RtError(rc, lib().clode_last_error().decode(errors="replace"))


def device_count() -> int:
    n = ctypes.c_int(0)
    _check(lib().clode_device_count(ctypes.byref(n)))
    return n.value


def device_info(device: int = 0) -> DeviceInfoC:
    info = DeviceInfoC()
    _check(lib().clode_device_get_info(device, ctypes.byref(info)))
    return info


def measure_fp64_peak(device: int = 0, repeats: int = 5) -> tuple[float, float]:
    """(TFLOP/s, best ms) of a register-resident DFMA micro-benchmark on `device`"""
    tf, ms = ctypes.c_double(), ctypes.c_double()
    _check(lib().clode_measure_fp64_peak(device, repeats, ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value


@dataclass
class Program:
    """Description of one JIT-specialised program (clode_program_desc)."""
    rhs_source: str
    stepper: str
    n_var: int
    n_par: int
    n_aux: int = 0
    n_wiener: int = 0
    observer: str = "basic"
    single_precision: bool = False
    f_var_ix: int = 0
    e_var_ix: int = 0
    n_store_events: int = 0
    kernels: int = KERNEL_TRANSIENT | KERNEL_FEATURES | KERNEL_TRAJECTORY
    bit_exact: bool = False
    work_queue: bool = False
    block_size: int = 0
    min_blocks_per_sm: int = 0
    staged_trajectory: bool = False
    observer_in_shared: bool = False
    ieee_constant_division: bool = False
    library_exp: bool = False

    def c(self) -> ProgramDesc:
        return ProgramDesc(self.rhs_source.encode(), self.stepper.encode(), self.observer.encode(),
                           int(self.single_precision), self.n_var, self.n_par, self.n_aux, self.n_wiener,
                           self.f_var_ix, self.e_var_ix, self.n_store_events, self.kernels,
                           int(self.bit_exact), int(self.work_queue), self.block_size, self.min_blocks_per_sm,
                           int(self.staged_trajectory), int(self.observer_in_shared),
                           int(self.ieee_constant_division), int(self.library_exp))


class _PinnedBlock:
    """a clode_host_alloc block, freed with the last array that views it"""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib().clode_host_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


class PinnedArray(np.ndarray):
    """ndarray over page-locked host memory; `_block` (reached through .base by every view) owns the memory"""
    _block = None


def pinned_empty(count: int, device: int = 0) -> np.ndarray:
    """flat float64 array in page-locked host memory (clode_host_alloc): device<->host copies run at the full
    PCIe rate and, for the streamed trajectory, overlap the integration"""
    count = int(count)
    f = lib().clode_host_alloc
    f.restype = ctypes.c_void_p
    f.argtypes = [ctypes.c_int, ctypes.c_size_t]
    lib().clode_host_free.argtypes = [ctypes.c_void_p]
    lib().clode_host_free.restype = None
    ptr = f(device, max(count, 1) * 8)
    if not ptr:
        raise RtError(-1, lib().clode_last_error().decode(errors="replace"))
    buf = (ctypes.c_double * max(count, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64, count=count).view(PinnedArray)
    arr._block = _PinnedBlock(ptr)
    return arr


def compile_program(prog: Program) -> tuple[bytes, str]:
    """NVRTC-compile to an sm_100a cubin; needs no GPU."""
    d = prog.c()
    cubin, size, log = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
    rc = lib().clode_compile(ctypes.byref(d), ctypes.byref(cubin), ctypes.byref(size), ctypes.byref(log))
    text = ctypes.string_at(log).decode(errors="replace") if log else ""
    if log:
        lib().clode_free(log)
    if rc:
        raise RtError(rc, lib().clode_last_error().decode(errors="replace"))
    data = ctypes.string_at(cubin, size.value)
    lib().clode_free(cubin)
    return data, text


def program_source(prog: Program) -> str:
    d = prog.c()
    src = ctypes.c_void_p()
    _check(lib().clode_program_source(ctypes.byref(d), ctypes.byref(src)))
    text = ctypes.string_at(src).decode(errors="replace")
    lib().clode_free(src)
    return text


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Sim:
    """One ensemble on one GPU (a `clode_sim`). Arrays are flat and variable-major."""

    def __init__(self, prog: Program, device: int = 0):
        self._h = ctypes.c_void_p()
        self._lib = lib()
        _check(self._lib.clode_sim_create(device, ctypes.byref(self._h)))
        self.prog = None
        self.n = 0
        self.build(prog)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.clode_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, prog: Program):
        d = prog.c()
        _check(self._lib.clode_sim_build(self._h, ctypes.byref(d)))
        self.prog = prog

    # ---- problem data ------------------------------------------------------------------
    def set_problem(self, x0, pars, dt: float | None = None):
        x0, pars = _f64(x0).ravel(), _f64(pars).ravel()
        n = x0.size // self.prog.n_var
        if n * self.prog.n_var != x0.size or n * self.prog.n_par != pars.size:
            raise ValueError("x0 / pars sizes do not describe the same number of instances")
        fill = self._sp.dt if dt is None and hasattr(self, "_sp") else (0.1 if dt is None else dt)
        _check(self._lib.clode_sim_set_npts(self._h, ctypes.c_size_t(n), ctypes.c_double(fill)))
        self.n = n
        self.set_x0(x0)
        self.set_pars(pars)

    def set_x0(self, x0):
        x0 = _f64(x0).ravel()
        _check(self._lib.clode_sim_set_x0(self._h, x0.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x0.size)))

    def set_pars(self, pars):
        pars = _f64(pars).ravel()
        _check(self._lib.clode_sim_set_pars(self._h, pars.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(pars.size)))

    def set_dt(self, dt):
        dt = _f64(dt).ravel()
        _check(self._lib.clode_sim_set_dt(self._h, dt.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(dt.size)))

    def set_tspan(self, t0, t1):
        _check(self._lib.clode_sim_set_tspan(self._h, ctypes.c_double(t0), ctypes.c_double(t1)))

    def set_solver_params(self, dt=0.1, dtmax=0.5, abstol=1e-6, reltol=1e-3, max_steps=1000000,
                          max_store=1000000, nout=1):
        self._sp = SolverParamsC(dt, dtmax, abstol, reltol, max_steps, max_store, nout)
        _check(self._lib.clode_sim_set_solver_params(self._h, ctypes.byref(self._sp)))

    def set_observer_params(self, e_var_ix=0, f_var_ix=0, max_event_count=100, max_event_timestamps=0,
                            min_amp=0.0, min_imi=0.0, nhood_radius=0.05, x_up_threshold=0.2,
                            x_down_threshold=0.2, dx_up_threshold=0.0, dx_down_threshold=0.0, eps_dx=0.0):
        op = ObserverParamsC(e_var_ix, f_var_ix, max_event_count, max_event_timestamps, min_amp, min_imi,
                             nhood_radius, x_up_threshold, x_down_threshold, dx_up_threshold,
                             dx_down_threshold, eps_dx)
        _check(self._lib.clode_sim_set_observer_params(self._h, ctypes.byref(op)))

    def seed_rng(self, seed: int, offset: int = 0, n_global: int = 0):
        _check(self._lib.clode_sim_seed_rng(self._h, ctypes.c_int64(seed), ctypes.c_uint64(offset),
                                            ctypes.c_uint64(n_global)))

    def set_rng_state(self, state):
        state = np.ascontiguousarray(state, dtype=np.uint64).ravel()
        _check(self._lib.clode_sim_set_rng_state(self._h, state.ctypes.data_as(ctypes.c_void_p),
                                                 ctypes.c_size_t(state.size)))

    def get_rng_state(self):
        out = np.empty(2 * self.n, np.uint64)
        _check(self._lib.clode_sim_get_rng_state(self._h, out.ctypes.data_as(ctypes.c_void_p),
                                                 ctypes.c_size_t(out.size)))
        return out

    # ---- simulation --------------------------------------------------------------------
    def transient(self):
        _check(self._lib.clode_sim_transient(self._h))

    def initialize_observer(self):
        _check(self._lib.clode_sim_initialize_observer(self._h))

    def features(self, initialize: int = -1):
        _check(self._lib.clode_sim_features(self._h, int(initialize)))

    def trajectory(self):
        _check(self._lib.clode_sim_trajectory(self._h))

    def shift_x0(self):
        _check(self._lib.clode_sim_shift_x0(self._h))

    # ---- results -----------------------------------------------------------------------
    def get(self, which: int, count: int, out: np.ndarray | None = None):
        """download `count` reals of buffer `which`; pass a preallocated (ideally pinned) float64 `out` to avoid
        a pageable bounce buffer on large transfers"""
        if out is None:
            out = np.empty(count, np.float64)
        elif out.dtype != np.float64 or out.size < count or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array with at least `count` elements")
        _check(self._lib.clode_sim_get(self._h, which, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(count)))
        return out[:count] if out.size != count else out

    def set_rows(self, which: int, host: np.ndarray, rows: int, pitch: int, first: int = 0, stride: int = 1):
        """clode_sim_set_rows: element (row r, local column k) is host.flat[r*pitch + first + k*stride]"""
        host = np.asarray(host)
        assert host.dtype == np.float64 and host.flags.c_contiguous
        _check(self._lib.clode_sim_set_rows(self._h, which, host.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(rows),
                                            ctypes.c_size_t(pitch), ctypes.c_size_t(first), ctypes.c_size_t(stride)))

    def get_rows(self, which: int, host: np.ndarray, rows: int, pitch: int, first: int = 0, stride: int = 1):
        assert host.dtype == np.float64 and host.flags.c_contiguous
        _check(self._lib.clode_sim_get_rows(self._h, which, host.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(rows),
                                            ctypes.c_size_t(pitch), ctypes.c_size_t(first), ctypes.c_size_t(stride)))
        return host

    def get_x0(self): return self.get(BUF_X0, self.n * self.prog.n_var)
    def get_xf(self, out=None): return self.get(BUF_XF, self.n * self.prog.n_var, out)
    def get_dt(self): return self.get(BUF_DT, self.n)
    def get_tf(self): return self.get(BUF_TF, self.n)

    def n_features(self) -> int:
        k = ctypes.c_int()
        _check(self._lib.clode_sim_n_features(self._h, ctypes.byref(k)))
        return k.value

    def get_f(self, out=None): return self.get(BUF_F, self.n * self.n_features(), out)

    def get_trajectory(self):
        rows = self._sp.max_store + 1
        nv, na, n = self.prog.n_var, self.prog.n_aux, self.n
        nst = np.empty(n, np.int32)
        _check(self._lib.clode_sim_get_n_stored(self._h, nst.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n)))
        return dict(t=self.get(BUF_T, rows * n), x=self.get(BUF_X, rows * n * nv), dx=self.get(BUF_DX, rows * n * nv),
                    aux=self.get(BUF_AUX, rows * n * na) if na else np.zeros(1), n_stored=nst, rows=rows)

    def trajectory_stream(self, chunk_rows: int, out: dict | None = None, pinned: bool = False, want=("t", "x", "dx", "aux")):
        """clode_sim_trajectory_stream: integrate in chunks of `chunk_rows` stored points, copying each chunk to
        the host while the next one integrates.  Returns dict(t, x, dx, aux, n_stored, rows) with rows = max_store
        (the rows the reference API returns).  `out` may hold preallocated flat float64 arrays (e.g. from
        pinned_empty); outputs not listed in `want` are skipped."""
        rows = self._sp.max_store
        nv, na, n = self.prog.n_var, self.prog.n_aux, self.n
        sizes = dict(t=rows * n, x=rows * n * nv, dx=rows * n * nv, aux=rows * n * na)
        res, ptr = {}, {}
        for k, size in sizes.items():
            if k not in want or size == 0:
                res[k], ptr[k] = (np.zeros(1) if k == "aux" else None), None
                continue
            a = out.get(k) if out else None
            if a is None:
                a = pinned_empty(size, device=0) if pinned else np.zeros(size, np.float64)
            elif a.dtype != np.float64 or a.size != size or not a.flags.c_contiguous:
                raise ValueError(f"out[{k!r}] must be a C-contiguous float64 array of {size} elements")
            res[k], ptr[k] = a, a.ctypes.data_as(ctypes.c_void_p)
        nst = np.empty(n, np.int32)
        _check(self._lib.clode_sim_trajectory_stream(self._h, ctypes.c_size_t(chunk_rows), ptr["t"], ptr["x"], ptr["dx"],
                                                     ptr["aux"], nst.ctypes.data_as(ctypes.c_void_p)))
        res.update(n_stored=nst, rows=rows)
        return res

    def get_trajectory_counts(self):
        nst = np.empty(self.n, np.int32)
        _check(self._lib.clode_sim_get_n_stored(self._h, nst.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(self.n)))
        return nst

    def get_steps(self):
        out = np.empty(self.n, np.uint32)
        _check(self._lib.clode_sim_get_steps(self._h, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(self.n)))
        return out

    def device_buffer(self, which: int):
        ptr, nbytes, elem = ctypes.c_uint64(), ctypes.c_size_t(), ctypes.c_int()
        _check(self._lib.clode_sim_device_buffer(self._h, which, ctypes.byref(ptr), ctypes.byref(nbytes), ctypes.byref(elem)))
        return ptr.value, nbytes.value, elem.value

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float()
        _check(self._lib.clode_sim_last_kernel_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        k = ctypes.c_uint64()
        _check(self._lib.clode_sim_launch_count(self._h, ctypes.byref(k)))
        return k.value

    def kernel_info(self, kernel: int) -> dict:
        info = KernelInfoC()
        _check(self._lib.clode_sim_kernel_info(self._h, kernel, ctypes.byref(info)))
        return {n: getattr(info, n) for n, _ in KernelInfoC._fields_}


def gather_rows(sims, which: int, rows: int, n_total: int, out: np.ndarray | None = None) -> np.ndarray:
    """clode_gather_rows: interleaved shards `sims` (shard g = instances g, g+G, ...) -> host [rows][n_total] through one
    NVLink gather to sims[0]'s GPU and one device-to-host copy"""
    if out is None:
        out = pinned_empty(rows * n_total, device=0)
    handles = (ctypes.c_void_p * len(sims))(*[s._h for s in sims])
    _check(lib().clode_gather_rows(handles, len(sims), which, ctypes.c_size_t(rows), ctypes.c_size_t(n_total),
                                   out.ctypes.data_as(ctypes.c_void_p)))
    return out
