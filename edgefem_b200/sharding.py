"""Frequency-sweep sharding across ranks (one process per GPU).  The sweep has no data-path
collective: rank r owns frequencies r, r+W, r+2W, ... (round-robin keeps the per-rank Krylov
iteration counts balanced, since iterations grow with frequency) and the P x P S-matrices are
all-gathered once at the end (16*P^2 bytes per point; NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_indices(n_points: int, rank: int, world: int) -> np.ndarray:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return np.arange(rank, n_points, world, dtype=np.int64)


def gather_sweep(local_S: np.ndarray, n_points: int, rank: int, world: int, dist=None, device=None) -> np.ndarray:
    """All-gather the per-rank [n_local, P, P] complex S arrays into the full [n_points, P, P] sweep.
    `dist` is torch.distributed (initialised) or None for a single process."""
    local_S = np.asarray(local_S, dtype=np.complex128)
    P = local_S.shape[1] if local_S.ndim == 3 else 0
    if world == 1 or dist is None:
        out = np.zeros((n_points, P, P), dtype=np.complex128)
        out[shard_indices(n_points, 0, 1)] = local_S
        return out
    import torch

    n_max = (n_points + world - 1) // world  # ranks may own one point fewer: pad to a common length
    buf = np.zeros((n_max, P, P, 2), dtype=np.float64)
    buf[: local_S.shape[0], ..., 0] = local_S.real
    buf[: local_S.shape[0], ..., 1] = local_S.imag
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.zeros((n_points, P, P), dtype=np.complex128)
    for r in range(world):
        idx = shard_indices(n_points, r, world)
        a = parts[r].cpu().numpy()
        out[idx] = a[: idx.size, ..., 0] + 1j * a[: idx.size, ..., 1]
    return out


def row_range(m: int, rank: int, world: int):
    """Row block of rank `rank` in the row-partitioned single-system solve (same rule as efb_dist_row_range):
    chunk = ceil(m / world), rows [rank*chunk, min(m, (rank+1)*chunk)) -- every rank but the last owns exactly
    `chunk` rows, so the owner of a global row is row // chunk."""
    if m <= 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad m/rank/world")
    chunk = (m + world - 1) // world
    return min(m, rank * chunk), min(m, (rank + 1) * chunk)


def gather_rows(local_x: np.ndarray, m: int, rank: int, world: int, dist=None, device=None) -> np.ndarray:
    """All-gather the row blocks of a distributed vector into the full length-m vector (tests / post-processing)."""
    local_x = np.asarray(local_x, dtype=np.complex128)
    if world == 1 or dist is None:
        return local_x.copy()
    import torch

    chunk = (m + world - 1) // world
    buf = np.zeros((chunk, 2), dtype=np.float64)
    buf[: local_x.size, 0] = local_x.real
    buf[: local_x.size, 1] = local_x.imag
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.zeros(m, dtype=np.complex128)
    for r in range(world):
        a, b = row_range(m, r, world)
        p = parts[r].cpu().numpy()
        out[a:b] = p[: b - a, 0] + 1j * p[: b - a, 1]
    return out
