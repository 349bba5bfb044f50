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
