#!/usr/bin/env python
"""Timing probe of the persistent small-system solvers on the WR-90 fixture.

  python tools/cluster_probe.py --points 256 [--reps 3]

Prints one JSON line: kernel ms, launch shape (CTAs per cluster, rhs per job, resident clusters), iterations and
microseconds per matrix-iteration.  EDGEFEM_B200_CLUSTER=0|C and EDGEFEM_B200_CLUSTER_NR=1|2 select the variant."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from edgefem_b200 import cabi, load_pyedgefem  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    pe = load_pyedgefem()
    ctx = cabi.Ctx(0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "rect_waveguide.npz"))
    hm = pe.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], z["node_ids"].tolist())
    bc = pe.build_edge_pec(hm, 1)
    dims = pe.RectWaveguidePort(bench.WR90_A, bench.WR90_B)
    kc_sq = (math.pi / bench.WR90_A) ** 2
    ports = [pe.build_wave_port_2d(hm, tag, pe.solve_te10_mode(dims, 10e9), set(bc.dirichlet_edges), kc_sq) for tag in (2, 3)]
    freqs = list(np.linspace(bench.F_LO, bench.F_HI, a.points)) if a.points > 1 else [10e9]
    rs = bench.ResidentSweep(ctx, pe, hm, bc, ports, freqs)
    out = []
    for _ in range(a.reps):
        ctx.timer_start()
        S, res = rs.step()
        ms_step = ctx.timer_stop()
        ms = rs.sys.last_solve_kernel_ms()
        shape = rs.sys.last_solve_shape()
        iters = [r["iters"] for r in res]
        P = rs.P
        it_m = [max(iters[f * P:(f + 1) * P]) for f in range(rs.F)]
        out.append({"ms_step": ms_step, "ms_kernel": ms, "shape": shape, "matrix_iterations": int(sum(it_m)), "rhs_iterations": int(sum(iters)),
                    "converged": all(r["converged"] for r in res)})
    o = out[-1]
    c, nr, ncl = o["shape"]
    if c:
        jobs_it = o["rhs_iterations"] if nr == 1 else o["matrix_iterations"]
        o["us_per_job_iteration"] = 1000.0 * o["ms_kernel"] * min(ncl, len(freqs) * (2 // nr)) / jobs_it
    else:
        o["us_per_job_iteration"] = 1000.0 * o["ms_kernel"] * min(148, len(freqs)) / o["matrix_iterations"]
    print(json.dumps({"points": a.points, "runs": out}))


if __name__ == "__main__":
    main()
