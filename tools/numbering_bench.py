#!/usr/bin/env python
"""Edge numbering at scale: device (efb_build_edges) vs the C++ host walk.  python tools/numbering_bench.py [n]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import edgefem_b200  # noqa: E402
from edgefem_b200 import cabi, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pe = edgefem_b200.load_pyedgefem()
xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
ctx = cabi.Ctx(0)
cabi.build_edges_device(ctx, tets[:1000], tris[:10])  # warm up (module load, allocator)
t0 = time.perf_counter()
te, to, re_, ro, edges = cabi.build_edges_device(ctx, tets, tris)
t_dev = time.perf_counter() - t0
k_ms = ctx.last_kernel_ms()
t0 = time.perf_counter()
hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
t_host = time.perf_counter() - t0
ok = np.array_equal(te, hm.tet_edges_array()) and np.array_equal(edges, hm.edges_array()) and np.array_equal(re_, hm.tri_edges_array())
print({"n": n, "tets": int(tets.shape[0]), "edges": int(edges.shape[0]), "device_total_s": round(t_dev, 3), "device_kernels_ms": round(k_ms, 2),
       "host_mesh_from_arrays_s": round(t_host, 3), "bit_exact": bool(ok)})
