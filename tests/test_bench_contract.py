"""bench.py's host-side pieces that need no GPU: workload definitions, the reference arm's sampling of the job's global
grid (the CPU arm must integrate the same parameter range as the GPU arm at every N), the step-count
row of every observer."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


@pytest.mark.parametrize("name, n_par, n_var", [("C1", 1, 2), ("C2", 3, 3), ("C2l", 3, 3), ("C3", 3, 4), ("C4", 4, 4), ("C5", 3, 3), ("C5d", 3, 3)])
def test_workloads_are_well_formed(name, n_par, n_var):
    n = 1 << 10
    w = bench.workload(name, n, np.arange(n))
    assert w["pars"].size == n_par * n and w["x0"].size == n_var * n and w["flops_per_step"] > 0
    # a shard of the global grid sees the global parameter range (interleaved sharding, any N)
    for world in (2, 8):
        shard = bench.workload(name, n * world, np.arange(3, n * world, world))
        whole = bench.workload(name, n * world, np.arange(n * world))
        p_s, p_w = shard["pars"].reshape(n_par, -1), whole["pars"].reshape(n_par, -1)
        assert np.array_equal(p_s, p_w[:, 3::world])
        span = np.maximum(p_w.max(axis=1) - p_w.min(axis=1), 1e-12)
        assert np.all(np.abs(p_s.min(axis=1) - p_w.min(axis=1)) <= 0.05 * span + 1e-12)
        assert np.all(np.abs(p_s.max(axis=1) - p_w.max(axis=1)) <= 0.05 * span + 1e-12)


@pytest.mark.parametrize("n_gpus", [1, 2, 8])
def test_reference_arm_samples_the_jobs_global_grid(n_gpus):
    base, ms = bench.cpu_reference("C2", n_gpus, steps=1, warmup=0, sample=64)
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] > 1e5 and ms > 0
    assert f"of the {n_gpus}-GPU job" in base["sample"] and f"64 of {n_gpus << 20} instances" in base["sample"]


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["e2e"]["h2d_bytes_per_step"] == 0
    # every other rank exits without work
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, env=dict(os.environ, RANK="1"), timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_step_count_rows():
    assert bench.step_row("basic", 6) == 5 and bench.step_row("basicall", 24) == 23 and bench.step_row("localmax", 26) == 25
    assert bench.step_row("thresh2", 46) == 42 and bench.step_row("nhood2", 44) == 40
