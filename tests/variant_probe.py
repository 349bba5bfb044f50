"""Helper of test_gpu_variants.py: assembles a small jittered PEC cube with the kernel variant the environment selects
(EDGEFEM_B200_ASM_KERNEL, EDGEFEM_B200_ASM_NO_REAL, EDGEFEM_B200_SPMV_KERNEL are read once per process) and writes the
matrix values and one SpMV result to the .npz named on the command line."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from edgefem_b200 import cabi, load_pyedgefem, meshgen  # noqa: E402


def main(out, n=7):
    pe = load_pyedgefem()
    ctx = cabi.Ctx(0)
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
    tp = np.asarray(tp).copy()
    tp[::3] = 7  # a second material region so that two slots are exercised
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    bc = pe.build_edge_pec(hm, 1)
    dm = cabi.DeviceMesh(ctx, hm.xyz_array(), hm.tet_nodes_array(), hm.tet_edges_array(), hm.tet_orient_array(), hm.tet_phys_array(),
                         hm.edge_nodes_array())
    flags = np.zeros(hm.num_edges(), dtype=np.uint8)
    flags[np.asarray(bc.dirichlet_edges, dtype=np.int64)] = 1
    idx = np.nonzero(flags)[0].astype(np.int32)
    sysd = cabi.DeviceSystem.from_mesh(dm, idx, idx, n_matrix=1, n_rhs=1)
    sysd.set_dirichlet(flags)
    omega = (2 * math.pi / (10 * (1.0 / n))) * 299792458.0
    ns = len(dm.slot_tags)
    res = {}
    for name, eps in (("real", [1.0, 2.2][:ns] + [1.0] * max(0, ns - 2)), ("lossy", [1.0, 2.2 - 0.05j][:ns] + [1.0] * max(0, ns - 2))):
        mats, keep = cabi.make_materials(ns, eps=np.asarray(eps, dtype=np.complex128))
        sysd.assemble_volume([omega], mats)
        res["vals_" + name] = sysd.values(0)
        rng = np.random.default_rng(7)
        x = rng.standard_normal(sysd.m) + 1j * rng.standard_normal(sysd.m)
        res["y_" + name] = sysd.spmv(0, x)
    rp, ci = sysd.pattern()
    res["rowptr"], res["colidx"] = rp, ci
    np.savez(out, **res)
    sysd.close()
    dm.close()


if __name__ == "__main__":
    main(sys.argv[1])
