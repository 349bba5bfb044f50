#!/usr/bin/env python
"""Row-partitioned solve of the synthetic PEC cube over the GPUs of one node.

  python tools/dist_cube.py --cube-n 40                       # one GPU (world 1 exercises the same kernels)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/dist_cube.py --cube-n 64 --check

Every rank builds the whole mesh on the host, uploads it, creates ITS row block of the system, assembles it (no
communication) and takes part in the distributed COCG+Jacobi solve.  Prints one JSON line on rank 0: assembly ms,
distributed SpMV ms / GB/s (peer-load halo vs NCCL all-gather baseline), COCG iteration ms, solve iterations/residual,
and with --check the difference to the single-GPU solve of the same system.  Times are max over ranks, CUDA events.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C0 = 299792458.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube-n", dest="n", type=int, default=40)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-it", type=int, default=4000)
    ap.add_argument("--kh", type=float, default=0.2, help="k0*h of the solve (small = well conditioned, SPD-like)")
    ap.add_argument("--freq", type=float, default=0.0, help="if > 0: solve frequency in Hz (overrides --kh); 240e6 is the C5 solve of bench.py")
    ap.add_argument("--precond", default="both", choices=["jacobi", "aux", "both"])
    ap.add_argument("--check", action="store_true", help="also solve on one GPU (rank 0) and compare")
    a = ap.parse_args()
    import numpy as np

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("gloo")  # plumbing only: ships the NCCL id and gathers results; the solve uses its own communicator
    import edgefem_b200
    from edgefem_b200 import cabi, meshgen, sharding

    ctx = cabi.Ctx(local)
    uid = [cabi.dist_unique_id() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(uid, src=0)
    ctx.dist_init(rank, world, uid[0])

    t0 = time.perf_counter()
    n = a.n
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
    # large-mesh ingest: edges numbered on the GPU (bit-exact with build_edges), no per-element host Mesh
    dm, info = cabi.device_mesh_from_conn(ctx, xyz, tets, tp, tris)
    m = int(info["edges"].shape[0])
    flags = cabi.pec_flags_from_tris(m, info["tri_edges"], trp, 1)
    r0, r1 = cabi.dist_row_range(m, rank, world)
    assert (r0, r1) == sharding.row_range(m, rank, world)
    sysd = cabi.DeviceSystem.from_mesh_rows(dm, r0, r1)
    sysd.set_dirichlet(flags)
    setup_s = time.perf_counter() - t0
    omega = 2 * math.pi * a.freq if a.freq > 0 else a.kh * n * C0  # k0 = kh / h, h = 1/n
    mats, keep = cabi.make_materials(len(dm.slot_tags))
    sysd.assemble_volume([omega], mats)  # first call also builds the rank-major assembly schedule of this system
    ctx.timer_start()
    sysd.assemble_volume([omega], mats)
    ms_asm = ctx.timer_stop()
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    b[flags == 1] = 0
    sysd.rhs_set(0, b[r0:r1])

    def maxr(v):
        if dist is None:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(v):
        if dist is None:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    out = {"n": n, "world": world, "tets": int(tets.shape[0]), "edges": m, "nnz_total": int(sumr(sysd.nnz)), "rows_local": r1 - r0,
           "host_setup_s": round(setup_s, 2), "assembly_ms": maxr(ms_asm)}
    nnz_tot = out["nnz_total"]
    b_spmv = nnz_tot * 20.0 + m * 36.0
    b_cocg = b_spmv + 10 * 16.0 * m
    for mode, name in ((0, "peer_load"), (1, "nccl_allgather")):
        ms_spmv = maxr(sysd.dist_bench(0, a.reps, mode))
        ms_it = maxr(sysd.dist_bench(1, a.reps, mode))
        out[name] = {"spmv_ms": ms_spmv, "spmv_gbs_total": b_spmv / ms_spmv / 1e6, "cocg_iteration_ms": ms_it, "cocg_gbs_total": b_cocg / ms_it / 1e6}
    ms_aux = maxr(sysd.dist_bench(2, a.reps, 0))
    b_aux = b_cocg + 2 * 16.0 * m + 3 * 16.0 * int(xyz.shape[0]) + 2 * m * 8.0
    out["peer_load"]["cocg_aux_iteration_ms"] = ms_aux
    out["peer_load"]["cocg_aux_gbs_total"] = b_aux / ms_aux / 1e6
    sysd.rhs_set(0, b[r0:r1])
    x_loc = None
    if a.precond in ("jacobi", "both"):
        ctx.sync()
        t1 = time.perf_counter()
        res = sysd.dist_solve(tol=a.tol, max_iterations=a.max_it, halo_mode=0)
        out["solve"] = dict(res, wall_s=round(time.perf_counter() - t1, 3))
        x_loc = sysd.x_get(0)
    if a.precond in ("aux", "both"):
        ctx.sync()
        t1 = time.perf_counter()
        res_a = sysd.dist_solve(tol=a.tol, max_iterations=a.max_it, halo_mode=0, precond=cabi.PRECOND_AUX)
        wall = time.perf_counter() - t1
        out["solve_aux"] = dict(res_a, wall_s=round(wall, 3), ms_per_iteration=1e3 * wall / max(1, res_a["iters"]))
        x_aux = sysd.x_get(0)
        if x_loc is None:
            x_loc = x_aux
    if a.check:
        res1 = sysd.dist_solve(tol=a.tol, max_iterations=a.max_it, halo_mode=1)
        x_loc1 = sysd.x_get(0)
        out["solve_allgather"] = res1
        x = sharding.gather_rows(x_loc, m, rank, world, dist=dist)
        x1 = sharding.gather_rows(x_loc1, m, rank, world, dist=dist)
        out["halo_modes_rel_diff"] = float(np.linalg.norm(x - x1) / max(np.linalg.norm(x), 1e-300))
        xa = sharding.gather_rows(x_aux, m, rank, world, dist=dist) if a.precond in ("aux", "both") else None
        if rank == 0:
            # single-GPU reference of the same system through the ordinary path
            pe_idx = np.nonzero(flags)[0].astype(np.int32)
            s1 = cabi.DeviceSystem.from_mesh(dm, pe_idx, pe_idx, n_matrix=1, n_rhs=1)
            s1.set_dirichlet(flags)
            s1.assemble_volume([omega], mats)
            s1.rhs_set(0, b)
            r = s1.solve(method=cabi.METHOD_COCG, precond=cabi.PRECOND_JACOBI, tol=a.tol, max_iterations=a.max_it)[0]
            xs = s1.x_get(0)
            out["single_gpu"] = r
            out["rel_diff_vs_single_gpu"] = float(np.linalg.norm(x - xs) / np.linalg.norm(xs))
            # true residual of the distributed solution with the single-GPU matrix
            y = s1.spmv(0, x)
            out["true_residual_dist"] = float(np.linalg.norm(b - y) / np.linalg.norm(b))
            if xa is not None:
                out["rel_diff_aux_vs_single_gpu"] = float(np.linalg.norm(xa - xs) / np.linalg.norm(xs))
                out["true_residual_dist_aux"] = float(np.linalg.norm(b - s1.spmv(0, xa)) / np.linalg.norm(b))
            s1.close()
    if rank == 0:
        print(json.dumps(out))
    sysd.close()
    dm.close()
    if dist is not None:
        dist.barrier()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
