#!/usr/bin/env python
"""Synthetic PEC cube (BASELINE config 5 family): build, assemble, SpMV, Krylov iterations, optional solve.
Used for ncu captures and for the large-mesh throughput numbers:  python tools/prof_cube.py --n 64 [--solve]"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from edgefem_b200 import cabi, load_pyedgefem, meshgen  # noqa: E402

C0 = 299792458.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--solve", action="store_true")
    ap.add_argument("--max-it", type=int, default=2000)
    a = ap.parse_args()
    pe = load_pyedgefem()
    ctx = cabi.Ctx(0)
    t0 = time.perf_counter()
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(a.n, jitter=0.1)
    t1 = time.perf_counter()
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    bc = pe.build_edge_pec(hm, 1)
    t2 = time.perf_counter()
    dm = cabi.DeviceMesh(ctx, hm.xyz_array(), hm.tet_nodes_array(), hm.tet_edges_array(), hm.tet_orient_array(), hm.tet_phys_array(),
                         hm.edge_nodes_array())
    flags = np.zeros(hm.num_edges(), dtype=np.uint8)
    flags[np.asarray(bc.dirichlet_edges, dtype=np.int64)] = 1
    pe_idx = np.nonzero(flags)[0].astype(np.int32)
    sysd = cabi.DeviceSystem.from_mesh(dm, pe_idx, pe_idx, n_matrix=1, n_rhs=1)
    sysd.set_dirichlet(flags)
    t3 = time.perf_counter()
    h = 1.0 / a.n
    omega = (2 * math.pi / (10 * h)) * C0
    mats, keep = cabi.make_materials(len(dm.slot_tags))
    sysd.assemble_volume([omega], mats)
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(sysd.m) + 1j * rng.standard_normal(sysd.m)
    b[flags == 1] = 0
    sysd.rhs_set(0, b)
    sysd.x_set(0, b)
    n_tet, n_node, m, nnz = hm.num_tets(), hm.num_nodes(), sysd.m, sysd.nnz
    out = {"n": a.n, "tets": n_tet, "nodes": n_node, "edges": m, "nnz": nnz,
           "host_s": {"generate": t1 - t0, "edges+pec": t2 - t1, "upload+pattern": t3 - t2}}
    for name, which, byts in (("assembly", 3, 45.0 * n_tet + 24.0 * n_node + 16.0 * nnz), ("spmv", 0, nnz * 20.0 + m * 36.0),
                              ("bicgstab_jacobi_it", 1, 2 * (nnz * 20.0 + m * 36.0) + 21 * 16.0 * m),
                              ("cocg_jacobi_it", 2, nnz * 20.0 + m * 36.0 + 10 * 16.0 * m)):
        ms = sysd.bench_kernel(which, a.reps)
        out[name] = {"ms": ms, "gbs": byts / ms / 1e6, "algorithmic_gb": byts / 1e9}
    out["assembly"]["mtets_per_s"] = n_tet / out["assembly"]["ms"] / 1e3
    if a.solve:
        sysd.assemble_volume([omega], mats)
        sysd.rhs_set(0, b)
        t4 = time.perf_counter()
        res = sysd.solve(precond=cabi.PRECOND_AUX, tol=1e-8, max_iterations=a.max_it, symmetric=True)
        out["solve"] = dict(res[0], wall_s=time.perf_counter() - t4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
