#!/usr/bin/env python
"""Small solves of the persistent kernels for compute-sanitizer (memcheck / racecheck / synccheck):

  compute-sanitizer --tool racecheck python tools/sanitize_probe.py cluster 6    # cluster-split kernel, 6 CTAs per cluster, aux
  compute-sanitizer --tool memcheck  python tools/sanitize_probe.py small 90     # one-CTA kernel, 90 matrices (head split)
  (small 3: resident one-rhs jobs; small 200: two rounds)

Results of round 2: profiles/sanitizer_r02.md."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import edgefem_oracle as orc
from edgefem_b200 import cabi, meshgen
import helpers as H
ctx = cabi.Ctx(0)
xyz, tets, tp, tris, trp = meshgen.rect_waveguide(a=0.02286, b=0.01016, length=0.03, nx=6, ny=3, nz=10)
mesh = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
pec = orc.build_edge_pec(mesh, 1)
f = 10e9
ports = orc.wr90_ports(mesh, pec, f)
S_ref = orc.wr90_sparams(mesh, pec, f, ports)
mode = sys.argv[1]
if mode == "cluster":
    os.environ["EDGEFEM_B200_CLUSTER"] = sys.argv[2]
    S, res = H.eigenmode_sweep_gpu(ctx, mesh, pec, ports, [f], method=cabi.METHOD_COCG, precond=cabi.PRECOND_AUX)
else:
    os.environ["EDGEFEM_B200_CLUSTER"] = "0"
    n = int(sys.argv[2])
    S, res = H.eigenmode_sweep_gpu(ctx, mesh, pec, ports, list(np.linspace(9e9, 11e9, n)), method=cabi.METHOD_COCG, precond=cabi.PRECOND_AUX)
print(mode, sys.argv[2], "converged", all(r["converged"] for r in res), "err", float(np.max(np.abs(S[0] - S_ref))) if mode == "cluster" else "-")
