"""Row-partitioned single-system solve (efb_dist_*): world 1 in-process, world 2 through torchrun when two GPUs exist."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
C0 = 299792458.0


def _cube(pe, n):
    from edgefem_b200 import meshgen

    xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    bc = pe.build_edge_pec(hm, 1)
    flags = np.zeros(hm.num_edges(), dtype=np.uint8)
    flags[np.asarray(bc.dirichlet_edges, dtype=np.int64)] = 1
    return hm, flags


def test_row_block_assembly_and_world1_solve():
    """A row block assembled on its own equals the same rows of the whole-mesh assembly bit for bit, and the
    distributed solver (world 1: same kernels, NCCL communicator of one rank) solves the system."""
    import edgefem_b200
    from edgefem_b200 import cabi

    pe = edgefem_b200.load_pyedgefem()
    ctx = cabi.Ctx(0)
    hm, flags = _cube(pe, 6)
    m = hm.num_edges()
    dm = cabi.DeviceMesh(ctx, hm.xyz_array(), hm.tet_nodes_array(), hm.tet_edges_array(), hm.tet_orient_array(), hm.tet_phys_array(),
                         hm.edge_nodes_array())
    mats, keep = cabi.make_materials(len(dm.slot_tags))
    omega = 0.2 * 6 * C0
    pe_idx = np.nonzero(flags)[0].astype(np.int32)
    whole = cabi.DeviceSystem.from_mesh(dm, pe_idx, pe_idx)
    whole.set_dirichlet(flags)
    whole.assemble_volume([omega], mats)
    rp, ci = whole.pattern()
    vals = whole.values(0)
    # two row blocks (as ranks 0 and 1 of a world of 2 would own them)
    for r in range(2):
        a, b = cabi.dist_row_range(m, r, 2)
        blk = cabi.DeviceSystem.from_mesh_rows(dm, a, b)
        blk.set_dirichlet(flags)
        blk.assemble_volume([omega], mats)
        brp, bci = blk.pattern()
        assert np.array_equal(brp, rp[a:b + 1] - rp[a])
        assert np.array_equal(bci, ci[rp[a]:rp[b]])
        assert np.array_equal(blk.values(0), vals[rp[a]:rp[b]])  # same kernel, same summation order: bit-exact
        with pytest.raises(cabi.EfbError):
            blk.solve()  # global-id entry points refuse row blocks
        blk.close()
    # world-1 distributed solve vs SuperLU on the assembled matrix
    ctx.dist_init(0, 1, cabi.dist_unique_id())
    one = cabi.DeviceSystem.from_mesh_rows(dm, 0, m)
    one.set_dirichlet(flags)
    one.assemble_volume([omega], mats)
    rng = np.random.default_rng(7)
    bvec = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    bvec[flags == 1] = 0
    one.rhs_set(0, bvec)
    A = sp.csr_matrix((vals, ci, rp), shape=(m, m))
    x_ref = spla.splu(A.tocsc()).solve(bvec)
    for mode in (0, 1):
        res = one.dist_solve(tol=1e-11, max_iterations=20000, halo_mode=mode)
        assert res["converged"], res
        x = one.x_get(0)
        assert np.linalg.norm(A @ x - bvec) / np.linalg.norm(bvec) < 1e-10
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-8
    assert one.dist_bench(0, 3, 0) > 0 and one.dist_bench(1, 3, 1) > 0
    # auxiliary-space preconditioner on the row partition: same solution, several times fewer iterations than Jacobi
    it_j = res["iters"]
    for mode in (0, 1):
        res = one.dist_solve(tol=1e-11, max_iterations=20000, halo_mode=mode, precond=cabi.PRECOND_AUX)
        assert res["converged"] and res["precond"] == cabi.PRECOND_AUX, res
        x = one.x_get(0)
        assert np.linalg.norm(A @ x - bvec) / np.linalg.norm(bvec) < 1e-10
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-8
        assert res["iters"] < 0.7 * it_j, (res["iters"], it_j)
    assert one.dist_bench(2, 3, 0) > 0
    one.close()
    whole.close()
    dm.close()
    ctx.close()


def test_world2_torchrun_matches_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29533", os.path.join(ROOT, "tools", "dist_cube.py"), "--cube-n", "12", "--check", "--tol", "1e-10", "--reps", "3", "--max-it", "20000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["solve"]["converged"] and d["solve_allgather"]["converged"]
    assert d["true_residual_dist"] < 1e-9
    assert d["rel_diff_vs_single_gpu"] < 1e-7
    assert d["halo_modes_rel_diff"] < 1e-9
    # the same system with the auxiliary-space preconditioner distributed over the two ranks, against SuperLU-quality single-GPU AUX
    assert d["solve_aux"]["converged"] and d["solve_aux"]["iters"] < d["solve"]["iters"]
    assert d["rel_diff_aux_vs_single_gpu"] < 1e-7
