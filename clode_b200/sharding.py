"""Host-side logic of the one-process-per-GPU path: instance-range partition, slicing of the
variable-major ensemble arrays, and the single exchange step — the final gather of per-shard result
matrices to rank 0 (NCCL over NVLink on GPU tensors, gloo on CPU tensors in the tests).

Instances are independent, so nothing else is communicated.  The in-process
multi-GPU path of the C++ classes (CLODE::setNpts) uses the INTERLEAVED assignment (`interleaved` below: shard g
owns instances g, g+G, ...); `partition` (contiguous ranges) is kept for callers that need contiguous slices.

Seeding follows CLODE::seedRNG(cl_int) (clode/cpp/CLODE.cpp:447-453) literally: the word index is added to the seed
in 32-bit signed arithmetic (`mySeed + (cl_int)i` wraps) and the sum is sign-extended to 64 bits — `_seed_words`.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def partition(n_total: int, world: int, rank: int, align: int = 32) -> Tuple[int, int]:
    """[lo, hi) of `rank`: contiguous chunks of ceil(n/world) rounded up to `align` (the last may be short or empty)"""
    chunk = -(-n_total // world)
    chunk = -(-chunk // align) * align
    lo = min(rank * chunk, n_total)
    return lo, min(lo + chunk, n_total)


def interleaved(n_total: int, world: int, rank: int) -> np.ndarray:
    """Cost-balanced assignment: rank g owns instances g, g+world, g+2*world, ...  Parameter sweeps are
    usually sorted grids whose cost varies smoothly along the grid (Lorenz r-sweep: 300 steps at r<1, 6500 in
    the chaotic band), so contiguous ranges load the GPUs unevenly while every interleaved shard sees the
    same cost distribution.  On its own GPU a shard is still stored contiguously (coalesced)."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def take_rows(flat: np.ndarray, rows: int, n_total: int, index: np.ndarray) -> np.ndarray:
    """[rows][n_total] flat -> [rows][len(index)] flat for an arbitrary instance index set"""
    return np.ascontiguousarray(np.asarray(flat).reshape(rows, n_total)[:, index]).ravel()


def _seed_words(seed: int, k: np.ndarray) -> np.ndarray:
    """RNGstate[k] = (cl_ulong)(mySeed + (cl_int)k): int32 wrap-around, then sign extension"""
    s = (np.int64(seed) + np.asarray(k, dtype=np.int64)).astype(np.int32)  # wraps modulo 2^32
    return s.astype(np.int64).astype(np.uint64)


def seed_states_for(seed: int, n_total: int, index: np.ndarray) -> np.ndarray:
    """global seeding rule (clode/cpp/CLODE.cpp:447-453) for an arbitrary instance index set"""
    i = np.asarray(index, dtype=np.int64)
    return np.concatenate([_seed_words(seed, i), _seed_words(seed, np.int64(n_total) + i)])


def gather_interleaved(local, rows: int, n_total: int, dst: int = 0):
    """gather_rows for interleaved shards: rank g's column k is global instance g + k*world"""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    widest = -(-n_total // world)
    mine = len(range(rank, n_total, world))
    padded = torch.zeros(rows * widest, dtype=local.dtype, device=local.device)
    if mine:
        padded.view(rows, widest)[:, :mine] = local.view(rows, mine)
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, parts, dst=dst)
    if rank != dst:
        return None
    out = torch.empty(rows, n_total, dtype=local.dtype, device=local.device)
    for g, part in enumerate(parts):
        cnt = len(range(g, n_total, world))
        if cnt:
            out[:, g::world] = part.view(rows, widest)[:, :cnt]
    return out.reshape(-1)


def shard_rows(flat: np.ndarray, rows: int, n_total: int, lo: int, hi: int) -> np.ndarray:
    """[rows][n_total] flat -> [rows][hi-lo] flat"""
    return np.ascontiguousarray(np.asarray(flat).reshape(rows, n_total)[:, lo:hi]).ravel()


def seed_states(seed: int, n_total: int, lo: int, hi: int) -> np.ndarray:
    """RNG state words of instances [lo, hi) under the reference's global rule RNGstate[k] = seed + k
    (clode/cpp/CLODE.cpp:447-453): instance i owns words i and n_total + i"""
    return seed_states_for(seed, n_total, np.arange(lo, hi, dtype=np.int64))


def gather_rows(local, rows: int, n_total: int, dst: int = 0):
    """Gather per-rank [rows][n_local] tensors into [rows][n_total] on `dst` (None elsewhere).
    `local` is a 1-D torch tensor (CPU for gloo, CUDA for NCCL); ranks may hold different n_local."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [partition(n_total, world, r) for r in range(world)]
    widest = max(hi - lo for lo, hi in counts)
    lo, hi = counts[rank]
    padded = torch.zeros(rows * widest, dtype=local.dtype, device=local.device)
    if hi > lo:
        padded.view(rows, widest)[:, : hi - lo] = local.view(rows, hi - lo)
    parts: Optional[List] = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, parts, dst=dst)
    if rank != dst:
        return None
    out = torch.empty(rows, n_total, dtype=local.dtype, device=local.device)
    for (plo, phi), part in zip(counts, parts):
        if phi > plo:
            out[:, plo:phi] = part.view(rows, widest)[:, : phi - plo]
    return out.reshape(-1)


class PipelinedGather:
    """The exchange step of the one-process-per-GPU path, overlapped with the next pass: `submit(local)` copies this rank's
    result matrix into one of two staging tensors and starts an ASYNCHRONOUS gather to rank `dst` (NCCL over NVLink for CUDA
    tensors, gloo for CPU tensors in the tests); the caller goes on with the next pass while the bytes move.  A staging
    tensor is reused only after its previous gather has completed; `drain()` completes what is in flight and returns the
    last gathered matrix in the global interleaved layout ([rows][n_total] flat) on `dst`, None elsewhere.
    Ranks hold interleaved shards (rank g: instances g, g+world, ...)."""

    def __init__(self, rows: int, n_total: int, dst: int = 0):
        import torch.distributed as dist

        self.rows, self.n_total, self.dst = rows, n_total, dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.widest = -(-n_total // self.world)
        self.mine = len(range(self.rank, n_total, self.world))
        self.stage = [None, None]
        self.parts = [None, None]
        self.work = [None, None]
        self.k = 0
        self.last = None

    def submit(self, local):
        import torch
        import torch.distributed as dist

        b = self.k & 1
        self.k += 1
        if self.work[b] is not None:
            self.work[b].wait()
            self.work[b] = None
        if self.stage[b] is None:
            self.stage[b] = torch.zeros(self.rows * self.widest, dtype=local.dtype, device=local.device)
            if self.rank == self.dst:
                self.parts[b] = [torch.empty_like(self.stage[b]) for _ in range(self.world)]
        if self.mine:
            self.stage[b].view(self.rows, self.widest)[:, :self.mine].copy_(local.view(self.rows, self.mine))
        self.work[b] = dist.gather(self.stage[b], self.parts[b], dst=self.dst, async_op=True)
        self.last = b

    def drain(self):
        import torch

        for b in (0, 1):
            if self.work[b] is not None:
                self.work[b].wait()
                self.work[b] = None
        if self.last is None or self.rank != self.dst:
            return None
        parts = self.parts[self.last]
        out = torch.empty(self.rows, self.n_total, dtype=parts[0].dtype, device=parts[0].device)
        for g, part in enumerate(parts):
            cnt = len(range(g, self.n_total, self.world))
            if cnt:
                out[:, g::self.world] = part.view(self.rows, self.widest)[:, :cnt]
        return out.reshape(-1)
