#!/usr/bin/env python
"""Stage trace of one C3 frequency point (normalize_port_weights + calculate_sparams): EDGEFEM_B200_TRACE=2 python tools/patch_trace.py"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgefem_b200 import meshgen, load_pyedgefem  # noqa: E402

pe = load_pyedgefem()
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.7
xyz, tets, tp, tris, trp, info = meshgen.patch_antenna(hmax_scale=scale)
hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
bc = pe.BC()
for tag in (1, 2, 3, 4, 10):
    bc.merge(pe.build_edge_pec(hm, tag))
opts = pe.SolveOptions()
opts.use_direct = True
f = 2.45e9
ph = pe.MaxwellParams()
ph.omega = 2 * math.pi * f
ph.use_abc = True
ph.abc_surface_tags = {50}
ph.set_eps_r_region(110, complex(4.4, -0.088))
cfg = pe.LumpedPortConfig()
cfg.surface_tag, cfg.z0, cfg.e_direction = 5, 50.0, [1.0, 0, 0]
for rep in range(2):
    t = time.time()
    ports = pe.normalize_port_weights(hm, ph, bc, [pe.build_lumped_port(hm, cfg)], opts)
    t1 = time.time()
    S = pe.calculate_sparams(hm, ph, bc, ports, opts)
    print("tets", hm.num_tets(), "edges", hm.num_edges(), "normalize %.3f s, calculate %.3f s, |S11| %.4f" % (t1 - t, time.time() - t1, abs(S[0][0])), flush=True)
