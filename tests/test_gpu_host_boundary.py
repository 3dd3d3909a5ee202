"""The host boundary of the runtime: strided and transposing uploads through the
page-locked staging ring, result arrays in pooled page-locked memory, the NVLink gather of interleaved shards."""
import numpy as np
import pytest

from oracle.common import MODELS
from problems import ensemble, rhs_source

pytestmark = pytest.mark.gpu


def _sim(rt, model="lorenz63", stepper="dopri5", observer="basic", device=0, single=False):
    nv, npar, na, nw = MODELS[model]
    prog = rt.Program(rhs_source(model), stepper, nv, npar, na, nw, observer=observer, kernels=rt.KERNEL_FEATURES,
                      single_precision=single)
    sim = rt.Sim(prog, device=device)
    sim.set_solver_params(dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, max_steps=1000000)
    sim.set_tspan(0.0, 5.0)
    return sim


@pytest.mark.parametrize("single", [False, True])
def test_cuda_strided_and_transposed_uploads(rt, single):
    n, G = 3 * 1024 * 1024 // 8 + 77, 3  # more than one 4 MiB staging chunk per row
    rng = np.random.default_rng(0)
    x0_matrix = rng.standard_normal((n * G, 3))          # the front end's (ensemble, nVar) layout
    flat = np.ascontiguousarray(x0_matrix.T).ravel()     # the API's variable-major layout, global ensemble
    sim = _sim(rt, single=single)
    sim.set_problem(np.zeros(3 * n), np.zeros(3 * n))
    want_all = flat.reshape(3, n * G)
    cast = (lambda a: a.astype(np.float32).astype(np.float64)) if single else (lambda a: a)
    for g in range(G):
        want = cast(want_all[:, g::G])
        sim.set_rows(rt.BUF_X0, flat, 3, n * G, first=g, stride=G)                 # shard g of a variable-major array
        assert np.array_equal(sim.get_x0().reshape(3, n), want)
        sim.set_x0(np.zeros(3 * n))
        sim.set_rows(rt.BUF_X0, x0_matrix.ravel(), 3, 1, first=g * 3, stride=G * 3)  # ... of an instance-major matrix
        assert np.array_equal(sim.get_x0().reshape(3, n), want)
        sim.set_x0(np.zeros(3 * n))
        rt._check(sim._lib.clode_sim_set_records(sim._h, rt.BUF_X0, x0_matrix.ctypes.data_as(__import__("ctypes").c_void_p),
                                                 __import__("ctypes").c_size_t(3), __import__("ctypes").c_size_t(3),
                                                 __import__("ctypes").c_size_t(g), __import__("ctypes").c_size_t(G)))  # records, transposed on the GPU
        assert np.array_equal(sim.get_x0().reshape(3, n), want)
        back = np.full(3 * n * G, np.nan)
        sim.get_rows(rt.BUF_X0, back, 3, n * G, first=g, stride=G)
        assert np.array_equal(back.reshape(3, n * G)[:, g::G], want)
        assert np.isnan(back.reshape(3, n * G)[:, (g + 1) % G::G]).all()           # other shards' columns untouched
    sim.close()


def test_cuda_gather_of_interleaved_shards_matches_one_gpu(rt):
    """three shards (on one GPU here; on their own GPUs in test_in_process_multi_gpu_*) integrate the interleaved
    thirds of an ensemble; clode_gather_rows assembles F and xf exactly as the unsharded run produces them"""
    n_total, G = 5000, 3
    _, x0, pars = ensemble("lorenz63", n_total)
    whole = _sim(rt)
    whole.set_problem(x0, pars)
    whole.features(1)
    F, xf = whole.get_f(), whole.get_xf()
    shards = []
    for g in range(G):
        s = _sim(rt)
        cnt = len(range(g, n_total, G))
        s.set_problem(np.zeros(3 * cnt), np.zeros(3 * cnt))
        s.set_rows(rt.BUF_X0, x0, 3, n_total, first=g, stride=G)
        s.set_rows(rt.BUF_PARS, pars, 3, n_total, first=g, stride=G)
        shards.append(s)
    for s in shards:  # all enqueued before any is waited for; the gather orders itself behind them with events
        rt._check(s._lib.clode_sim_enqueue(s._h, rt.KERNEL_FEATURES, 1))
    got_F = rt.gather_rows(shards, rt.BUF_F, 6, n_total)
    got_xf = rt.gather_rows(shards, rt.BUF_XF, 3, n_total, out=np.empty(3 * n_total))
    assert np.array_equal(got_F, F) and np.array_equal(got_xf, xf)
    for s in shards + [whole]:
        s.close()


def test_frontend_matrix_path_equals_flat_path_and_results_are_page_locked_views(rt):
    import clode_b200 as clode
    from clode_b200 import build
    from clode_b200.models import rhs_path
    from clode_b200.solver import _as_f

    build.build_all()
    n = 4096
    r = np.linspace(0.5, 60.0, n)

    def make():
        fs = clode.FeatureSimulator(src_file=rhs_path("lorenz63"), variables={"x": 1.0, "y": 1.0, "z": 1.0},
                                    parameters={"r": 28.0, "s": 10.0, "b": 8.0 / 3.0}, aux=["a"], observer=clode.Observer.basic,
                                    stepper=clode.Stepper.dormand_prince, single_precision=False, t_span=(0.0, 10.0), dt=0.01,
                                    dtmax=1.0, abstol=1e-6, reltol=1e-6)
        fs.set_ensemble(parameters={"r": r})
        return fs

    a = make()
    out = a.features()
    Fa = a._device_features                       # (n, nf) matrix, transposed on the GPU
    assert Fa.shape == (n, 6) and Fa.flags.c_contiguous and not Fa.flags.owndata
    assert np.shares_memory(out.F, Fa)            # the record array is a view of it, not a copy
    assert np.array_equal(np.ravel(out.get_var_max("x")), Fa[:, 0])
    b = make()                                    # the same data through the reference's flat (variable-major) calls
    b._integrator.set_problem_data(_as_f(b._device_initial_state), _as_f(b._device_parameters))
    b._integrator.features()
    Fb = np.asarray(b._integrator.get_f_array()).reshape((n, 6), order="F")
    assert np.array_equal(Fa, Fb)
    assert np.array_equal(a.get_final_state(), np.asarray(b._integrator.get_xf_array()).reshape((n, 3), order="F"))
    assert np.array_equal(np.asarray(b._integrator.get_pars_array()).reshape((n, 3), order="F")[:, 0], r)


def test_result_arrays_outlive_their_simulator(rt):
    """a page-locked result block keeps its own reference on the device context: closing the last simulation object
    (which releases the primary context) must not free memory that a numpy array — or the pool — still holds"""
    import gc

    import clode_b200 as clode
    from clode_b200 import build
    from clode_b200.models import rhs_path

    build.build_all()
    kept = []
    for k in range(3):
        fs = clode.FeatureSimulator(src_file=rhs_path("lorenz63"), variables={"x": 1.0, "y": 1.0, "z": 1.0},
                                    parameters={"r": 28.0 + k, "s": 10.0, "b": 8.0 / 3.0}, aux=["a"], observer=clode.Observer.basic,
                                    stepper=clode.Stepper.rk4, single_precision=(k == 1), t_span=(0.0, 1.0), dt=0.01)
        out = fs.features()
        kept.append((np.array(out.F["max x"]).copy(), out.F, fs.get_final_state()))
        del fs, out
        gc.collect()  # the simulator (and with it the last clode_sim) is gone; the arrays are not
    for snapshot, F, xf in kept:
        assert np.array_equal(F["max x"], snapshot) and np.all(np.isfinite(xf))
    a = rt.pinned_empty(1000)
    a[:] = 1.0
    del kept
    gc.collect()
    assert a.sum() == 1000.0


@pytest.mark.parametrize("single", [False, True])
def test_cuda_contiguous_chunks_plus_peer_pull_equals_strided_upload(rt, single):
    """the multi-GPU upload: contiguous chunk h of the host matrix to shard h (clode_sim_stage_records), then every shard pulls
    its interleaved instances out of all chunks (clode_scatter_records; peer loads over NVLink when the shards sit on
    different GPUs — on one GPU here, on two in test_in_process_*) — same device contents as the strided per-shard upload"""
    import ctypes

    n_total, G = 100003, 3  # ragged: the last chunk is shorter, the shards differ in size
    rng = np.random.default_rng(1)
    m = rng.standard_normal((n_total, 3))
    shards = []
    for g in range(G):
        s = _sim(rt, single=single)
        cnt = len(range(g, n_total, G))
        s.set_problem(np.zeros(3 * cnt), np.zeros(3 * cnt))
        shards.append(s)
    chunk = -(-n_total // G)
    for h, s in enumerate(shards):
        part = np.ascontiguousarray(m[h * chunk:(h + 1) * chunk])
        rt._check(s._lib.clode_sim_stage_records(s._h, part.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(part.shape[0]), ctypes.c_size_t(3)))
    handles = (ctypes.c_void_p * G)(*[s._h for s in shards])
    rt._check(rt.lib().clode_scatter_records(handles, G, rt.BUF_X0, ctypes.c_size_t(3), ctypes.c_size_t(n_total)))
    cast = (lambda a: a.astype(np.float32).astype(np.float64)) if single else (lambda a: a)
    for g, s in enumerate(shards):
        assert np.array_equal(s.get_x0().reshape(3, -1), cast(m[g::G].T))
    # staging is consumed: a second scatter without new chunks is a state error, not stale data
    assert rt.lib().clode_scatter_records(handles, G, rt.BUF_X0, ctypes.c_size_t(3), ctypes.c_size_t(n_total)) != 0
    for s in shards:
        s.close()
