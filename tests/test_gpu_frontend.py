"""The reference's own test-suite, replayed through this package's Python front end
(clode_b200.Simulator / FeatureSimulator / TrajectorySimulator over the pybind module and the C++
CLODE* classes) on the GPU.  Each test cites the reference test it mirrors.  Also covers in-process
sharding over several runtime objects and the C++-layer continuation semantics."""
import os
from math import exp, log, pi, sqrt

import numpy as np
import pytest

from oracle import restate
from oracle.common import Config, Observer as OObserver, Solver, seed_states
from problems import MODELS_DIR, ensemble
from util import assert_bit_equal, run_oracle

pytestmark = pytest.mark.gpu


def model(name):
    return os.path.join(MODELS_DIR, name + ".cl")


@pytest.fixture(scope="module")
def clode(rt):
    import clode_b200
    from clode_b200 import build

    build.build_all()
    return clode_b200


def test_ornl_thompson_a1(clode):
    """test/test_ornl_thompson_a1.py:16-46"""
    H = 10.0
    sim = clode.TrajectorySimulator(src_file=model("thompson_a1"), variables={"y1": 0.0, "y2": 0.0},
                                    parameters={"m": 0.25, "w": 8.0, "k": 2.0, "H": H}, aux=["g1"], num_noise=0,
                                    dt=0.001, dtmax=0.001, stepper=clode.Stepper.rk4, t_span=(0.0, H / 2.0),
                                    max_store=20000, max_steps=20000)
    tr = sim.trajectory()
    assert len(tr.t) >= 5000
    for tt, y1, y2 in zip(tr.t[1:], tr.x["y1"][1:], tr.x["y2"][1:]):
        np.testing.assert_approx_equal(y1, 4.0 * (tt + exp(-8.0 * tt) / 8.0 - 1.0 / 8.0), significant=5)
        np.testing.assert_approx_equal(y2, 4.0 * (1.0 - exp(-8.0 * tt)), significant=5)
    assert np.allclose(tr.aux["g1"], tr.x["y1"] - H, atol=1e-4)


def test_sine_curve_timestamps(clode):
    """test/test_features.py:24-87"""
    fs = clode.FeatureSimulator(src_file=model("sine_drive"), variables={"x": 0}, parameters={"dilation": 1},
                                aux=["xp1", "pos", "neg"], observer=clode.Observer.threshold_2,
                                stepper=clode.Stepper.rk4, dtmax=0.001, dt=0.001, t_span=(0.0, 4 * pi), event_var="x",
                                feature_var="x", observer_min_x_amp=0.5, observer_x_up_thresh=(2 + sqrt(2)) / 4,
                                observer_x_down_thresh=0.001, observer_dx_down_thresh=0.001,
                                observer_dx_up_thresh=0.001, observer_max_event_count=100,
                                observer_max_event_timestamps=3)
    out = fs.features()
    assert int(out.get_var_count("event")) == 2
    up, down = out.get_event_data("up"), out.get_event_data("down")
    assert len(up) == 2 and len(down) == 2
    assert np.isclose(up[0], pi / 4, atol=0.01) and np.isclose(up[1], 9 * pi / 4, atol=0.01)
    assert np.isclose(down[0], 3 / 2 * pi, atol=0.01) and np.isclose(down[1], 7 / 2 * pi, atol=0.01)
    assert np.array_equal(out.get_timestamps("up"), up) and np.array_equal(out.get_timestamps("down"), down)


@pytest.mark.parametrize("observer", ["basic_all_variables", "local_max", "neighbourhood_1", "neighbourhood_2", "threshold_2"])
def test_aux_values(clode, observer):
    """test/test_aux_values.py:27-91 (each observer actually used, unlike the reference's parametrisation)"""
    fs = clode.FeatureSimulator(src_file=model("sine_drive"), variables={"x": 0.0}, parameters={"dilation": 1.0},
                                aux=["xp1", "pos", "neg"], observer=getattr(clode.Observer, observer),
                                stepper=clode.Stepper.rk4, dtmax=0.1, dt=0.1, t_span=(0.0, 400 * pi), event_var="x",
                                feature_var="x")
    out = fs.features()
    assert out.get_var_mean("xp1") == pytest.approx(1.0, abs=1e-2)
    assert out.get_var_min("xp1") == pytest.approx(0.0, abs=1e-2)
    assert out.get_var_max("xp1") == pytest.approx(2.0, abs=1e-2)
    assert out.get_var_mean("pos") == 1.0 and out.get_var_min("pos") == 1.0 and out.get_var_max("pos") == 1.0
    assert out.get_var_min("neg") == -2.0 and out.get_var_max("neg") == -2.0


def _vdp_period(mu):
    if mu < 0:
        return 0
    if mu < 2:
        return 2 * pi * (1 + mu ** 2 / 16)
    return min(2 * pi * (1 + mu ** 2 / 16), (3 - 2 * log(2)) * mu + 3 * 2.2338 / mu ** (1 / 3.0))


def test_vdp_dormand_prince(clode):
    """test/test_vdp.py:50-95 — defaults include single precision"""
    integ = clode.FeatureSimulator(src_file=model("vanderpol"), variables={"x": 1.0, "y": 1.0}, parameters={"mu": 1.0},
                                   observer=clode.Observer.threshold_2, stepper=clode.Stepper.dormand_prince,
                                   t_span=(0.0, 1000.0), max_store=20000, max_steps=20000)
    mus = [-1, 0, 0.01, 0.1, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 5, 6]
    integ.set_ensemble(parameters={"mu": mus})
    integ.features()
    periods = integ.get_observer_results().get_var_max("period")
    for k, mu in enumerate(mus):
        assert np.isclose(periods[k], _vdp_period(mu), rtol=0.01, atol=1), (mu, periods[k])


def test_print_devices_and_log_levels(clode, capfd):
    """test/test_logger.py:4-28"""
    tr = clode.TrajectorySimulator(src_file=model("vanderpol"), variables={"x": 0.0, "y": 1.0}, parameters={"mu": 1.0},
                                   num_noise=0, stepper=clode.Stepper.dormand_prince, device_id=0, platform_id=0)
    clode.set_log_level(clode.LogLevel.trace)
    tr.print_devices()
    captured = capfd.readouterr()
    assert "OpenCL" in captured.out and "B200" in captured.out and captured.err == ""
    clode.set_log_level(clode.LogLevel.off)
    tr.print_devices()
    captured = capfd.readouterr()
    assert captured.out == "" and captured.err == ""
    clode.set_log_level(clode.LogLevel.warn)
    assert clode.query_opencl()[0].device_count >= 1


def test_transient_then_features_continuation_matches_oracle(clode):
    """the standard workflow transient() -> features() (examples/spike_counting.py:82-83) through the whole
    stack, double precision, compared with the oracle driven the same way (tolerance tier: FMA + libdevice)"""
    n = 64
    ts, x0, pars = ensemble("lactotroph", n)
    fs = clode.FeatureSimulator(src_file=model("lactotroph"), variables={"v": -60.0, "n": 0.0, "f": 0.0, "c": 0.1},
                                parameters={"gcal": 1.5, "gsk": 3.0, "gbk": 1.0}, aux=["ical"], single_precision=False,
                                stepper=clode.Stepper.rk4, dt=0.05, t_span=(0.0, 500.0),
                                observer=clode.Observer.threshold_2, event_var="v", feature_var="v")
    fs.set_ensemble(parameters=pars.reshape(3, n).T.copy())
    fs.transient()
    out = fs.features()
    lib = restate.OracleLib(Config("lactotroph", "rk4", "thresh2"))
    sp = Solver(dt=0.05, dtmax=1.0, max_steps=10000000)
    op = OObserver(x_up_threshold=0.3, x_down_threshold=0.2)
    r1 = lib.transient((0.0, 500.0), x0, pars, sp, np.full(n, sp.dt), seed_states(1, n))
    r2 = lib.features((0.0, 500.0), r1["xf"], pars, sp, op, r1["dt"], r1["rng"])
    F = out.to_ndarray()
    G = r2["F"].reshape(-1, n).T
    assert np.array_equal(F[:, 41], G[:, 41])  # event counts
    assert np.array_equal(F[:, 42], G[:, 42])  # step counts
    assert np.allclose(F, G, rtol=1e-6, atol=1e-8)
    assert np.allclose(fs.get_final_state(), r2["xf"].reshape(4, n).T, rtol=1e-7, atol=1e-9)


def test_in_process_sharding_over_runtime_objects(clode):
    """`device_ids=[...]` shards the ensemble over one runtime object per entry; results must not depend on it.
    (On a single-GPU box both shards live on device 0, which still exercises split / seed / gather.)"""
    n = 300
    ts, x0, pars = ensemble("lactotroph_noise", n)
    kw = dict(src_file=model("lactotroph_noise"), variables={"v": -60.0, "n": 0.0, "f": 0.0, "c": 0.1},
              parameters={"gcal": 1.5, "gsk": 3.0, "gbk": 1.0, "noise": 1.0}, aux=["ical"], num_noise=1,
              single_precision=False, stepper=clode.Stepper.stochastic_euler, dt=0.01, t_span=(0.0, 10.0),
              observer=clode.Observer.basic_all_variables, platform_id=0)
    results = []
    for ids in ([0], [0, 0], [0, 0, 0]):
        fs = clode.FeatureSimulator(device_ids=ids, **kw)
        fs.set_repeat_ensemble(n)
        fs.seed_rng(1)
        out = fs.features()
        results.append((out.to_ndarray(), fs.get_final_state(), fs._integrator.get_rng_state()))
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a, b)
    # and the stream itself is the reference's: compare with the oracle
    lib = restate.OracleLib(Config("lactotroph_noise", "seuler", "basicall"))
    r = lib.features((0.0, 10.0), x0, pars, Solver(dt=0.01, dtmax=1.0, max_steps=10000000), OObserver(),
                     np.full(n, 0.01), seed_states(1, n))
    assert np.array_equal(results[0][2], r["rng"])


def test_trajectory_store_limit_and_shapes(clode):
    """max_store cut-off: the kernel writes row index max_store; the API returns max_store rows"""
    sim = clode.TrajectorySimulator(src_file=model("chay_keizer"), variables={"v": -50.0, "n": 0.01, "c": 0.12},
                                    parameters={"gca": 800.0, "gkca": 750.0, "kpmca": 0.12}, single_precision=False,
                                    stepper=clode.Stepper.rk4, dt=0.5, t_span=(0.0, 1000.0), max_store=40, nout=2)
    sim.set_ensemble(parameters={"gca": np.linspace(550.0, 1050.0, 70)})
    res = sim.trajectory()
    assert len(res) == 70 and all(len(r.t) == 40 for r in res)
    assert np.allclose(res[3].t, np.arange(40) * 1.0)
    assert np.array_equal(np.asarray(sim._integrator.get_n_stored()), np.full(70, 40))


def test_python_and_xpp_right_hand_sides(clode, tmp_path):
    """test/test_vdp.py:33-95 passes the Van der Pol system as a Python function (`rhs_equation=`), and
    test/test_xpp_parser.py:21 as an XPP file; both must give exactly what the hand-written .cl gives"""
    src = ("def vdp(t: float, var: List[float], par: List[float], derivatives: List[float], aux: List[float], "
           "wiener: List[float]) -> None:\n    mu: float = par[0]\n    x: float = var[0]\n    y: float = var[1]\n"
           "    dx: float = y\n    dy: float = mu * (1 - x * x) * y - x\n    derivatives[0] = dx\n    derivatives[1] = dy\n")
    mod = tmp_path / "vdp_module.py"
    mod.write_text("from typing import List\n\n\n" + src)
    import importlib.util
    spec = importlib.util.spec_from_file_location("vdp_module", mod)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    xpp = tmp_path / "vdp.xpp"
    xpp.write_text("init x=1.0\ninit y=1.0\npar mu=1.0\ny' = mu * (1 - x*x) * y - x\nx' = y\n")
    kw = dict(variables={"x": 1.0, "y": 1.0}, parameters={"mu": 1.0}, observer=clode.Observer.threshold_2,
              stepper=clode.Stepper.dormand_prince, t_span=(0.0, 200.0), single_precision=False)
    mus = [0.01, 0.5, 1.0, 2.0, 4.0]
    results = []
    for source in (dict(src_file=model("vanderpol")), dict(rhs_equation=m.vdp), dict(src_file=str(xpp))):
        integ = clode.FeatureSimulator(**source, **kw)
        integ.set_ensemble(parameters={"mu": mus})
        integ.features()
        results.append((integ.get_observer_results().to_ndarray(), integ.get_final_state()))
    for other in results[1:]:
        assert np.allclose(results[0][0], other[0], rtol=1e-9, atol=1e-12, equal_nan=True)
        assert np.allclose(results[0][1], other[1], rtol=1e-9, atol=1e-12)


def test_streamed_trajectory_through_the_front_end(clode):
    """`stream_chunk_rows`: chunked launches with overlapped copy-out, on one runtime object and on
    in-process shards (device_ids=[0, 0, 0]); the TrajectoryOutput objects are identical to the single-launch ones."""
    kw = dict(src_file=model("chay_keizer"), variables={"v": -50.0, "n": 0.01, "c": 0.12},
              parameters={"gca": 800.0, "gkca": 750.0, "kpmca": 0.12}, single_precision=False,
              stepper=clode.Stepper.rk4, dt=0.5, t_span=(0.0, 300.0), max_store=400, nout=2, platform_id=0)
    gca = np.linspace(550.0, 1050.0, 70)

    def run(**extra):
        sim = clode.TrajectorySimulator(**kw, **extra)
        sim.set_ensemble(parameters={"gca": gca})
        res = sim.trajectory()
        return res, sim.get_final_state(), np.asarray(sim._integrator.get_n_stored())

    base = run(device_ids=[0])
    for extra in (dict(device_ids=[0], stream_chunk_rows=16), dict(device_ids=[0, 0, 0], stream_chunk_rows=64),
                  dict(device_ids=[0], stream_chunk_rows=100000)):
        other = run(**extra)
        assert np.array_equal(base[2], other[2]) and np.array_equal(base[1], other[1])
        for a, b in zip(base[0], other[0]):
            assert np.array_equal(a.t, b.t) and np.array_equal(a.to_ndarray("x"), b.to_ndarray("x"))
            assert np.array_equal(a.to_ndarray("dx"), b.to_ndarray("dx"))
    # switching back to the single launch on the same object
    sim = clode.TrajectorySimulator(**kw, device_ids=[0], stream_chunk_rows=8)
    sim.set_ensemble(parameters={"gca": gca})
    a = sim.trajectory(update_x0=False)
    sim.set_stream_chunk(0)
    b = sim.trajectory(update_x0=False)
    assert all(np.array_equal(p.t, q.t) and np.array_equal(p.to_ndarray("x"), q.to_ndarray("x")) for p, q in zip(a, b))


def test_in_process_two_gpus(clode, rt):
    """device_ids=[0, 1]: one runtime object per physical GPU, launches overlap, results identical to one GPU"""
    if rt.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 1 << 18  # enough blocks to fill two GPUs, so that the concurrency shows in the kernel time
    ts, x0, pars = ensemble("lorenz63", n)
    kw = dict(src_file=model("lorenz63"), variables={"x": 1.0, "y": 1.0, "z": 1.0},
              parameters={"r": 28.0, "s": 10.0, "b": 8.0 / 3.0}, aux=["dx"], single_precision=False,
              stepper=clode.Stepper.dormand_prince, dt=0.01, dtmax=1.0, abstol=1e-6, reltol=1e-6, t_span=(0.0, 30.0),
              observer=clode.Observer.local_max, platform_id=0)
    out = []
    for ids in ([0], [0, 1], [1]):
        fs = clode.FeatureSimulator(device_ids=ids, **kw)
        fs.set_ensemble(parameters=pars.reshape(3, n).T.copy())
        res = fs.features()
        out.append((res.to_ndarray(), fs.get_final_state(), fs._integrator.get_last_kernel_ms()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][0], out[2][0])
    # two GPUs working concurrently: the slower shard takes clearly less than the whole job on one GPU
    assert out[1][2] < 0.75 * out[0][2]
