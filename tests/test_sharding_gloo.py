"""World-size-2 gloo test of the multi-rank sweep plumbing (frequency sharding + S all-gather). CPU only."""
import os
import socket
import sys

import numpy as np
import pytest

from edgefem_b200 import sharding


def test_shard_indices_partition():
    for n, w in ((256, 1), (256, 8), (10, 4), (3, 4)):
        allidx = np.concatenate([sharding.shard_indices(n, r, w) for r in range(w)])
        assert sorted(allidx.tolist()) == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def _worker(rank, world, port, n_points, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sharding.shard_indices(n_points, rank, world)
    # a synthetic, rank-independent "S(f)" so every rank can verify the gathered result
    local = np.stack([np.array([[k + 1j, 2.0 * k], [3.0 - k * 1j, k * k]]) for k in idx]) if idx.size else np.zeros((0, 2, 2), complex)
    full = sharding.gather_sweep(local, n_points, rank, world, dist=dist)
    want = np.stack([np.array([[k + 1j, 2.0 * k], [3.0 - k * 1j, k * k]]) for k in range(n_points)])
    q.put((rank, bool(np.array_equal(full, want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_points", [7, 8])
def test_gather_sweep_world2(n_points):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_points, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_row_range_partition_matches_cabi():
    """Row blocks of the row-partitioned solve: disjoint cover, owner = row // chunk, same rule in the C library."""
    from edgefem_b200 import cabi

    for m, w in ((5745, 1), (5745, 2), (5745, 8), (10, 4), (9, 8), (23827950, 8)):
        chunk = (m + w - 1) // w
        prev = 0
        for r in range(w):
            a, b = sharding.row_range(m, r, w)
            assert a == prev and b >= a
            assert (a, b) == cabi.dist_row_range(m, r, w)
            for row in {a, b - 1} if b > a else ():
                assert row // chunk == r
            prev = b
        assert prev == m
    with pytest.raises(ValueError):
        sharding.row_range(10, 2, 2)


def _rows_worker(rank, world, port, m, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.row_range(m, rank, world)
    want = np.arange(m) * (1.0 + 0.5j) - 3.0j
    full = sharding.gather_rows(want[a:b], m, rank, world, dist=dist)
    q.put((rank, bool(np.array_equal(full, want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [11, 12])
def test_gather_rows_world2(m):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
