#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): refined PEC cube cavity, single frequency, assembly + converged Krylov solve.

  python tools/c5_solve.py --n 150 [--freq 240e6] [--precond aux|jacobi] [--tol 1e-10] [--max-it 40000]

The cube is the reference's cavity test object (tests/test_cavity_eigenmodes.cpp: 1 m PEC cube, modes at 212.0 MHz x3,
259.6 MHz x2, ...) refined to n^3 x 6 tets; the default frequency 240 MHz sits between the first two resonances.
Prints one JSON line."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from edgefem_b200 import cabi, meshgen  # noqa: E402

C0 = 299792458.0


def build(ctx, n):
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
    dm, info = cabi.device_mesh_from_conn(ctx, xyz, tets, tp, tris)
    flags = cabi.pec_flags_from_tris(info["edges"].shape[0], info["tri_edges"], trp, 1)
    sysd = cabi.DeviceSystem.from_mesh(dm, n_matrix=1, n_rhs=1)
    sysd.set_dirichlet(flags)
    return dm, sysd, flags, tets.shape[0], xyz.shape[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--freq", type=float, default=240e6)
    ap.add_argument("--kh", type=float, default=0.0, help="if > 0: omega such that k0 h = 2 pi / kh (overrides --freq)")
    ap.add_argument("--precond", default="aux")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--max-it", type=int, default=40000)
    ap.add_argument("--check-every", type=int, default=0)
    a = ap.parse_args()
    ctx = cabi.Ctx(0)
    t0 = time.perf_counter()
    dm, sysd, flags, n_tet, n_node = build(ctx, a.n)
    t_setup = time.perf_counter() - t0
    h = 1.0 / a.n
    omega = (2 * math.pi / (a.kh * h)) * C0 if a.kh > 0 else 2 * math.pi * a.freq
    mats, keep = cabi.make_materials(len(dm.slot_tags))
    sysd.assemble_volume([omega], mats)
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(sysd.m) + 1j * rng.standard_normal(sysd.m)
    b[flags == 1] = 0
    sysd.rhs_set(0, b)
    ctx.sync()
    t1 = time.perf_counter()
    pre = cabi.PRECOND_AUX if a.precond == "aux" else cabi.PRECOND_JACOBI
    res = sysd.solve(precond=pre, tol=a.tol, max_iterations=a.max_it, symmetric=True, check_every=a.check_every)
    ctx.sync()
    wall = time.perf_counter() - t1
    r = res[0]
    x = sysd.x_get(0)
    # independent check of the true residual on the device SpMV
    y = sysd.spmv(0, x)
    true_res = float(np.linalg.norm(b - y) / np.linalg.norm(b))
    m, nnz = sysd.m, sysd.nnz
    b_spmv = nnz * 20.0 + m * 36.0
    out = {"n": a.n, "tets": n_tet, "nodes": n_node, "edges": m, "nnz": nnz, "free_unknowns": int((flags == 0).sum()),
           "omega": omega, "k0h": omega / C0 * h, "setup_s": round(t_setup, 2), "precond": a.precond,
           "solve": {"iters": r["iters"], "converged": r["converged"], "residual": r["residual"], "true_residual_host_check": true_res,
                     "seconds": wall, "ms_per_iteration": 1000.0 * wall / max(1, r["iters"]),
                     "spmv_model_gbs_per_iteration": b_spmv / (wall / max(1, r["iters"])) / 1e9}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
