"""N > 1 host path on CPU: two processes (gloo), each integrates its shard of a stochastic ensemble
with the CPU oracle standing in for the GPU kernels, the feature matrices are gathered on rank 0 and
must equal the unsharded run bit for bit — which requires the global RNG seeding rule, the
variable-major slicing and the gather re-layout to be right.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clode_b200 import sharding
from oracle import restate
from oracle.common import Config, Observer, Solver
from problems import ensemble

N_TOTAL = 83  # not a multiple of anything


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _unsharded():
    lib = restate.OracleLib(Config("lactotroph_noise", "seuler", "basicall", math="pm"))
    ts, x0, pars = ensemble("lactotroph_noise", N_TOTAL)
    sp = Solver(dt=0.01, max_steps=100000)
    return lib.features((0.0, 3.0), x0, pars, sp, Observer(), np.full(N_TOTAL, sp.dt), sharding.seed_states(1, N_TOTAL, 0, N_TOTAL))


def _worker(rank, world, port, out_path, layout):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = restate.OracleLib(Config("lactotroph_noise", "seuler", "basicall", math="pm"))
    ts, x0, pars = ensemble("lactotroph_noise", N_TOTAL)
    if layout == "contiguous":
        lo, hi = sharding.partition(N_TOTAL, world, rank)
        index, gather = np.arange(lo, hi), sharding.gather_rows
    else:
        index, gather = sharding.interleaved(N_TOTAL, world, rank), sharding.gather_interleaved
    sp = Solver(dt=0.01, max_steps=100000)
    r = lib.features((0.0, 3.0), sharding.take_rows(x0, 4, N_TOTAL, index), sharding.take_rows(pars, 4, N_TOTAL, index),
                     sp, Observer(), np.full(index.size, sp.dt), sharding.seed_states_for(1, N_TOTAL, index))
    F = gather(torch.from_numpy(r["F"]), lib.n_feat, N_TOTAL)
    xf = gather(torch.from_numpy(r["xf"]), 4, N_TOTAL)
    rng = gather(torch.from_numpy(r["rng"].view(np.int64)), 2, N_TOTAL)
    if rank == 0:
        np.savez(out_path, F=F.numpy(), xf=xf.numpy(), rng=rng.numpy().view(np.uint64))
    dist.destroy_process_group()


@pytest.mark.parametrize("layout", ["contiguous", "interleaved"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_equals_unsharded(world, layout, tmp_path):
    restate.build(Config("lactotroph_noise", "seuler", "basicall", math="pm"))  # compile once, before forking
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), out, layout), nprocs=world, join=True)
    got, want = np.load(out), _unsharded()
    for k in ("F", "xf", "rng"):
        assert np.array_equal(got[k], want[k]), k


def test_partition_properties():
    for n, world in [(1 << 20, 8), (83, 2), (83, 3), (5, 8), (64, 2), (0, 4)]:
        spans = [sharding.partition(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % 32 == 0 or lo == n for lo, _ in spans)
    i = np.arange(10, 20)
    s = sharding.seed_states(-3, 100, 10, 20)
    assert np.array_equal(s[:10].astype(np.int64), -3 + i) and np.array_equal(s[10:].astype(np.int64), 97 + i)


def _pipelined_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, n_total = 3, 41
    g = sharding.PipelinedGather(rows, n_total)
    index = sharding.interleaved(n_total, world, rank)
    results = []
    for k in range(5):  # five passes; every staging tensor is reused while the other one's gather may be in flight
        full = np.arange(rows * n_total, dtype=np.float64).reshape(rows, n_total) * (k + 1)
        g.submit(torch.from_numpy(np.ascontiguousarray(full[:, index]).ravel()))
        if k in (1, 4):
            out = g.drain()
            if rank == 0:
                results.append((k, out.numpy().copy()))
    if rank == 0:
        np.savez(out_path, **{f"k{k}": v for k, v in results})
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pipelined_gather_delivers_the_last_submitted_pass(world, tmp_path):
    """bench.py's exchange step (asynchronous, double-buffered gather of interleaved shards): after drain() rank 0 holds
    the global matrix of the LAST submitted pass, whatever was in flight"""
    out = str(tmp_path / "pipelined.npz")
    mp.spawn(_pipelined_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    base = np.arange(3 * 41, dtype=np.float64)
    assert np.array_equal(got["k1"], base * 2) and np.array_equal(got["k4"], base * 5)


def test_seed_words_wrap_like_the_references_cl_int_arithmetic():
    """CLODE::seedRNG(cl_int): `RNGstate[i] = mySeed + i` is 32-bit signed (wraps), then widened to 64 bits with sign extension
    (clode/cpp/CLODE.cpp:447-453); the Python sharding helpers, the oracle helper and the C++ layer must agree near INT_MAX"""
    from oracle.common import seed_states as oracle_seed_states

    n = 10
    seed = 2**31 - 4
    got = sharding.seed_states_for(seed, n, np.arange(n))
    want = oracle_seed_states(seed, n)
    assert np.array_equal(got, want)
    k = np.arange(2 * n, dtype=np.int64)
    expect = ((seed + k + 2**31) % 2**32 - 2**31).astype(np.int64).astype(np.uint64)  # int32 wrap, sign extension
    assert np.array_equal(got, expect)
    assert got[3] == np.uint64(2**31 - 1) and got[4] == np.uint64(2**64 - 2**31)     # ... 0x7fffffff, then 0xffffffff80000000
    assert np.array_equal(sharding.seed_states(seed, n, 2, 7), sharding.seed_states_for(seed, n, np.arange(2, 7)))
