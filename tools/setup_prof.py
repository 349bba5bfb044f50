#!/usr/bin/env python
"""Stage times of the large-mesh set-up (array-only ingest).  python tools/setup_prof.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgefem_b200 import cabi, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = time.perf_counter
t = T()
xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
print("meshgen (test fixture, numpy)", round(T() - t, 3))
ctx = cabi.Ctx(0)
cabi.build_edges_device(ctx, tets[:1000], tris[:10])
t = T()
te, to, re_, ro, edges = cabi.build_edges_device(ctx, tets, tris)
t1 = T()
print("edge numbering (device)", round(t1 - t, 3))
tet_nodes, edge_nodes = (tets - 1).astype(np.int32), (edges - 1).astype(np.int32)
flags = cabi.pec_flags_from_tris(edges.shape[0], re_, trp, 1)
t2 = T()
print("index arrays + PEC flags (numpy)", round(t2 - t1, 3))
dm = cabi.DeviceMesh(ctx, xyz, tet_nodes, te, to, tp, edge_nodes)
t3 = T()
print("DeviceMesh", round(t3 - t2, 3))
s = cabi.DeviceSystem.from_mesh(dm)
t4 = T()
print("system create", round(t4 - t3, 3))
s.set_dirichlet(flags)
t5 = T()
print("set_dirichlet", round(t5 - t4, 3))
print("total set-up after mesh generation", round(t5 - t, 3), "s;", tets.shape[0], "tets", edges.shape[0], "edges", s.nnz, "nnz")
