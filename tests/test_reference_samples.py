"""Source-level drop-in check of the C++ boundary: the REFERENCE's own sample programs
(samples/testFeatures.cpp, samples/testTrajectory.cpp) are compiled UNMODIFIED — straight from the
reference tree, nothing copied — against this repository's CLODE / CLODEfeatures / CLODEtrajectory /
OpenCLResource headers and linked with libclode_rt.  The binaries land in oracle/_ref/samples/ (git-ignored,
travels to the GPU box), where the GPU half of this module runs them and compares their printed results
with the oracle."""
import glob
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from oracle import restate
from oracle.common import REFERENCE_ROOT, Config, Observer, Solver, seed_states

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
HOST = os.path.join(REPO, "clode_b200", "csrc", "host")
OUT = os.path.join(REPO, "oracle", "_ref", "samples")
SAMPLES = ["testFeatures", "testTrajectory"]


def test_reference_samples_compile_unmodified_against_our_headers(rt):
    src_dir = os.path.join(REFERENCE_ROOT, "samples")
    if not os.path.isdir(src_dir):
        pytest.skip("reference tree absent")
    os.makedirs(OUT, exist_ok=True)
    host_sources = [p for p in glob.glob(os.path.join(HOST, "*.cpp")) if not p.endswith("CLODEpython.cpp")]
    for name in SAMPLES:
        exe = os.path.join(OUT, name)
        cmd = ["g++", "-O2", "-std=c++17", "-w", f"-I{HOST}", f"-I{REPO}/include", os.path.join(src_dir, name + ".cpp"),
               *host_sources, "-o", exe, f"-L{REPO}/clode_b200", "-lclode_rt", f"-Wl,-rpath,{REPO}/clode_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        assert os.path.exists(exe)


def _run_sample(name, tmp_path):
    exe = os.path.join(OUT, name)
    if not os.path.exists(exe):
        pytest.skip("sample binary not built (needs the reference tree at build time)")
    os.makedirs(tmp_path / "samples", exist_ok=True)
    # the samples open "samples/lactotroph.cl"; our model file is arithmetic-identical (tests/test_oracle_pinning.py)
    shutil.copy(os.path.join(REPO, "clode_b200", "models", "lactotroph.cl"), tmp_path / "samples" / "lactotroph.cl")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REPO, "clode_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ERROR" not in r.stdout
    return r.stdout


@pytest.mark.gpu
def test_reference_testFeatures_sample_runs_and_matches_oracle(tmp_path):
    """samples/testFeatures.cpp:26-152 — lactotroph, rk4 dt=0.1, float, 32 instances, transient then thresh2 features"""
    out = _run_sample("testFeatures", tmp_path)
    feats = dict(re.findall(r"^ (.+?) = ([-+0-9.eE]+|nan|inf)$", out, flags=re.M))
    assert float(feats["step count"]) == 10001
    n = 32
    lib = restate.OracleLib(Config("lactotroph", "rk4", "thresh2", single=True))
    sp = Solver(dt=0.1, dtmax=1.0, abstol=1e-6, reltol=1e-3, max_steps=10000000, max_store=10000000, nout=50)
    op = Observer(0, 0, 100, 0, 0.0, 0.0, 0.01, 0.3, 0.2, 0.0, 0.0, 1e-7)
    pars = np.concatenate([np.full(n, 1.5), np.full(n, 3.0), np.full(n, 1.0)])
    r1 = lib.transient((0.0, 1000.0), np.zeros(4 * n), pars, sp, np.full(n, sp.dt), seed_states(1, n))
    r2 = lib.features((0.0, 1000.0), r1["xf"], pars, sp, op, r1["dt"], r1["rng"])
    F = r2["F"].reshape(-1, n)[:, 0].astype(np.float64)
    names = ["max period", "min period", "mean period", "max peaks", "min peaks", "mean peaks"]
    for k, name in enumerate(names):
        assert float(feats[name]) == pytest.approx(F[k], rel=2e-3, abs=1e-3), name  # single precision, printed to 6 digits
    assert float(feats["event count"]) == F[18 + 20 + 3]
    assert float(feats["max v"]) == pytest.approx(F[18], rel=2e-3)


@pytest.mark.gpu
def test_reference_testTrajectory_sample_runs(tmp_path):
    """samples/testTrajectory.cpp — dopri5, double, 2 instances; prints the stored trajectory of instance 0"""
    out = _run_sample("testTrajectory", tmp_path)
    m = re.search(r"Timepoints stored: (\d+)", out)
    assert m and 50 < int(m.group(1)) < 400  # the reference's pasted run stored 129 (samples/test_outputs_cpp.md:148)
    rows = [list(map(float, ln.split())) for ln in out.splitlines() if re.match(r"^[-0-9.e+]+\s+[-0-9.e+]+ [-0-9.e+]+ [-0-9.e+]+ [-0-9.e+]+\s*$", ln)]
    assert len(rows) >= 50
    t = np.array([r[0] for r in rows])
    v = np.array([r[1] for r in rows])
    assert np.all(np.diff(t) > 0) and t[-1] <= 1000.0 + 1e-6 and -80 < v.min() < v.max() < 40
