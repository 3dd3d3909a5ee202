"""The C-ABI runtime library: loads without a GPU, exports every symbol include/clode_rt.h
declares, NVRTC-compiles programs for sm_100a without a GPU, and fails LOUDLY (no CPU fallback)
when asked to compute without a CUDA driver.  CPU only; no compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def declared_symbols():
    text = open(os.path.join(REPO, "include", "clode_rt.h")).read()
    return sorted(set(re.findall(r"CLODE_API\s+[\w\s\*]+?\b(clode_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ["clode_sim_create", "clode_sim_build", "clode_sim_transient", "clode_sim_features",
                 "clode_sim_initialize_observer", "clode_sim_trajectory", "clode_sim_shift_x0", "clode_sim_get",
                 "clode_sim_seed_rng", "clode_compile", "clode_last_error", "clode_device_count"]:
        assert must in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol(rt):
    lib = rt.lib()
    out = subprocess.run(["nm", "-D", "--defined-only", rt.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (clode_\w+)", out))
    missing = [n for n in declared_symbols() if n not in exported]
    assert not missing, missing
    for n in declared_symbols():
        assert getattr(lib, n) is not None


def test_library_has_no_link_time_cuda_dependency(rt):
    out = subprocess.run(["ldd", rt.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda" not in out and "libnvrtc" not in out and "libtorch" not in out


def test_header_is_plain_c():
    src = '#include "clode_rt.h"\nint main(void){ clode_program_desc d; (void)d; return sizeof(clode_solver_params) == 48 ? 0 : 1; }\n'
    exe = os.path.join(REPO, "oracle", "_build", "abi_c_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{REPO}/include", "-x", "c", "-", "-o", exe],
                   input=src, text=True, check=True)
    assert subprocess.run([exe]).returncode == 0


def _have_driver(rt):
    try:
        return rt.device_count() > 0
    except rt.RtError:
        return False


def test_compute_without_driver_fails_loudly(rt):
    if _have_driver(rt):
        pytest.skip("a CUDA driver is present")
    with pytest.raises(rt.RtError) as e:
        rt.device_count()
    assert e.value.code == 2 and "libcuda" in str(e.value)
    from problems import rhs_source
    with pytest.raises(rt.RtError) as e:
        rt.Sim(rt.Program(rhs_source("lorenz63"), "rk4", 3, 3, 1))
    assert e.value.code == 2


@pytest.mark.parametrize("stepper", ["euler", "heun", "rk4", "bs23", "dopri5"])
@pytest.mark.parametrize("observer", ["basic", "basicall", "localmax", "nhood1", "nhood2", "thresh2"])
def test_nvrtc_compiles_every_stepper_observer_pair(rt, stepper, observer, tmp_path):
    from problems import rhs_source
    prog = rt.Program(rhs_source("lorenz63"), stepper, 3, 3, 1, observer=observer, n_store_events=2)
    cubin, log = rt.compile_program(prog)
    assert cubin[:4] == b"\x7fELF"
    p = tmp_path / "k.cubin"
    p.write_bytes(cubin)
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", str(p)], capture_output=True, text=True).stdout
    for k in ("clode_transient", "clode_initialize_observer", "clode_features", "clode_trajectory"):
        assert k in usage
    listing = subprocess.run(["cuobjdump", "-lelf", str(p)], capture_output=True, text=True).stdout
    assert "sm_100a" in listing, listing
    # the user RHS must be inlined into the kernels: no separate device function in the cubin
    functions = re.findall(r"Function (\w+):", usage)
    assert sorted(functions) == sorted(["clode_transient", "clode_initialize_observer", "clode_features",
                                        "clode_trajectory", "clode_observer_layout", "clode_sched_histogram",
                                        "clode_sched_scan", "clode_sched_scatter", "clode_interleave_rows", "clode_records_to_rows", "clode_records_pull"]), functions


def test_nvrtc_other_variants(rt):
    from problems import rhs_source
    rt.compile_program(rt.Program(rhs_source("lactotroph_noise"), "seuler", 4, 4, 1, 1, observer="basicall"))
    rt.compile_program(rt.Program(rhs_source("vanderpol"), "dopri5", 2, 1, observer="thresh2", single_precision=True))
    rt.compile_program(rt.Program(rhs_source("lactotroph"), "bs23", 4, 3, 1, observer="thresh2", bit_exact=True))
    rt.compile_program(rt.Program(rhs_source("chay_keizer"), "rk4", 3, 3, work_queue=True, kernels=rt.KERNEL_TRAJECTORY))
    src = rt.program_source(rt.Program(rhs_source("lorenz63"), "rk4", 3, 3, 1))
    assert "getRHS" in src and "-DEXPLICIT_RK4" in src and "clode_transient" in src


def test_build_errors_are_reported_with_the_compiler_log(rt):
    with pytest.raises(rt.RtError) as e:
        rt.compile_program(rt.Program("void getRHS(const realtype t, const realtype x_[], const realtype p_[], "
                                      "realtype dx_[], realtype aux_[], const realtype w_[]) { dx_[0] = undefined_symbol; }",
                                      "rk4", 1, 1))
    assert e.value.code == 4 and "undefined_symbol" in str(e.value)
    with pytest.raises(rt.RtError) as e:
        rt.compile_program(rt.Program("", "rk45", 1, 1))
    assert e.value.code == 1 and "unknown stepper" in str(e.value)
    with pytest.raises(rt.RtError):
        rt.compile_program(rt.Program("", "rk4", 2, 1, observer="thresh2", f_var_ix=5))


def test_bit_exact_tier_can_be_forced_from_the_environment(monkeypatch):
    """CLODE_BIT_EXACT=1: programs built without the bit_exact field (the C++ classes, the Python front end) get the
    bit-exact tier — no contraction, portable math, no PTX pass; single-precision programs are left alone"""
    from clode_b200 import _rt
    from clode_b200.models import MODELS, rhs_source

    nv, npar, na, nw = MODELS["lactotroph"]
    prog = _rt.Program(rhs_source("lactotroph"), "bs23", nv, npar, na, nw, observer="thresh2", kernels=_rt.KERNEL_FEATURES, min_blocks_per_sm=4)
    assert "-DCLODE_BITEXACT" not in _rt.program_source(prog).splitlines()[0]
    monkeypatch.setenv("CLODE_BIT_EXACT", "1")
    head = _rt.program_source(prog).splitlines()[0]
    assert "-DCLODE_BITEXACT" in head and "--fmad=false" in head and "-DCLODE_EXP_2K" not in head
    cubin, log = _rt.compile_program(prog)
    assert cubin[:4] == b"\x7fELF" and "ptx pass" not in log
    import dataclasses

    single = _rt.program_source(dataclasses.replace(prog, single_precision=True)).splitlines()[0]
    assert "-DCLODE_BITEXACT" not in single


def test_extents_go_to_shared_memory_only_when_the_array_fits(monkeypatch):
    """the observer extents in shared memory (CLODE_EXT_SMEM, observers.cuh) share the 48 KiB of static shared memory with
    the tables of the production math (16 KiB exp, 6 KiB polar method): a program for which the sum does not fit keeps the
    extents in registers instead of failing in ptxas; where it fits, one variable's words are loaded together (CLODE_EXT_BATCH)"""
    from clode_b200 import _rt
    from clode_b200.models import MODELS, rhs_source

    monkeypatch.setenv("CLODE_EXT_SMEM", "1")

    def head(model, stepper, observer, **kw):
        nv, npar, na, nw = MODELS[model]
        prog = _rt.Program(rhs_source(model), stepper, nv, npar, na, nw, observer=observer, kernels=_rt.KERNEL_FEATURES, min_blocks_per_sm=4, **kw)
        cubin, _ = _rt.compile_program(prog)
        assert cubin[:4] == b"\x7fELF"
        return _rt.program_source(prog).splitlines()[0]

    # 27 rows x 128 threads x 8 B = 27 KiB: fits beside the exp table (C3) ...
    fits = head("lactotroph", "bs23", "thresh2")
    assert "-DCLODE_EXT_SMEM" in fits and "-DCLODE_EXT_BATCH" in fits
    # ... but not beside exp table + polar table (stochastic stepper): registers
    assert "-DCLODE_EXT_SMEM" not in head("lactotroph_noise", "seuler", "thresh2")
    # the bit-exact tier stages no tables
    assert "-DCLODE_EXT_SMEM" in head("lactotroph_noise", "seuler", "thresh2", bit_exact=True)
    monkeypatch.setenv("CLODE_EXT_BATCH", "0")
    assert "-DCLODE_EXT_BATCH" not in head("lactotroph", "bs23", "thresh2")


def test_staged_trajectory_stores_fall_back_when_the_tile_does_not_fit():
    """staged_trajectory=1 needs a double-buffered tile of (1 + 2 nVar + nAux) rows in static shared memory beside the exp
    table: a 12-variable system does not fit in 48 KiB and is built with the direct stores instead of failing in ptxas"""
    from clode_b200 import _rt

    head = "void getRHS(const realtype t, const realtype x_[], const realtype p_[], realtype dx_[], realtype aux_[], const realtype w_[]) {\n"
    small = _rt.Program(head + "    dx_[0] = -x_[0]; dx_[1] = exp(-x_[1]); dx_[2] = x_[0];\n}\n", "rk4", 3, 1, 0,
                        kernels=_rt.KERNEL_TRAJECTORY, staged_trajectory=True)
    big = _rt.Program(head + "".join(f"    dx_[{j}] = -p_[0] * x_[{j}] + exp(-x_[{(j + 1) % 12}]);\n" for j in range(12)) + "}\n", "rk4", 12, 1, 0,
                      kernels=_rt.KERNEL_TRAJECTORY, staged_trajectory=True)
    assert "-DCLODE_TRAJ_STAGED" in _rt.program_source(small).splitlines()[0]
    assert "-DCLODE_TRAJ_STAGED" not in _rt.program_source(big).splitlines()[0]
    for prog in (small, big):
        cubin, _ = _rt.compile_program(prog)
        assert cubin[:4] == b"\x7fELF"
