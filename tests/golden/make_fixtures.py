"""Regenerates tests/golden/*.npz from the reference's shipped example meshes.

Run in the build container only (needs /root/reference):  python tests/golden/make_fixtures.py
The .npz files hold the raw Gmsh content (node ids, coordinates, tet/tri connectivity and
physical tags) -- data fixtures, not source -- so the GPU box (which has no /root/reference)
can run the parity tests.  Known answers copied from the reference's docs live in
tests/golden/kat.json with their file:line provenance.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import edgefem_oracle as orc  # noqa: E402

REF = "/root/reference/examples"
MESHES = ["rect_waveguide", "cube_cavity", "mixed_cavity", "wr42_waveguide", "dielectric_slab", "coax_50ohm"]


def main():
    for name in MESHES:
        m = orc.load_gmsh_v2(os.path.join(REF, name + ".msh"), fast=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), node_ids=m.node_ids, xyz=m.xyz, tet_conn=m.tet_conn,
                            tet_phys=m.tet_phys, tri_conn=m.tri_conn, tri_phys=m.tri_phys)
        print(name, m.xyz.shape[0], "nodes", m.tet_conn.shape[0], "tets", m.tri_conn.shape[0], "tris", m.num_edges, "edges")
    kat = {
        "wr90_table": {
            "source": "docs/validation.md:22-27 (freq GHz, |S11|, |S21|, phase(S21) deg)",
            "rows": [[7.0, 0.019, 0.9996, -150.6], [8.0, 0.003, 0.9997, 80.8], [9.0, 0.010, 0.9994, -15.6],
                     [10.0, 0.013, 0.9992, -100.7], [11.0, 0.041, 0.9982, 179.9], [12.0, 0.053, 0.9973, 103.9]],
        },
        "alpha_sweep_10ghz": {
            "source": "docs/validation.md:51-57 (port_abc_scale, |S11|, |S21|)",
            "rows": [[0.8, 0.207, 0.978], [0.9, 0.094, 0.995], [1.0, 0.013, 0.999], [1.2, 0.187, 0.982], [1.5, 0.387, 0.922]],
        },
        "wr90_counts": {"source": "docs/validation.md:14 + SURVEY F6", "tets": 4227, "edges": 5745, "nnz": 85113, "pec_edges": 1437},
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)


if __name__ == "__main__":
    main()
