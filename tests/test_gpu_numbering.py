"""device edge numbering (efb_build_edges) is bit-exact with the reference's first-seen hash-map walk
(oracle build_edges, src/mesh_gmsh.cpp:104-146) -- fixtures, shuffled/sparse node ids, tri-only edges, empty input."""
import os

import numpy as np
import pytest

import edgefem_oracle as orc
from conftest import GOLDEN, load_fixture_mesh
from edgefem_b200 import cabi, meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Ctx(0)
    yield c
    c.close()


def check(ctx, om):
    te, to, re_, ro, edges = cabi.build_edges_device(ctx, om.tet_conn, om.tri_conn)
    assert np.array_equal(te, om.tet_edges) and np.array_equal(to, om.tet_orient)
    assert np.array_equal(re_, om.tri_edges) and np.array_equal(ro, om.tri_orient)
    assert np.array_equal(edges, om.edges)


@pytest.mark.parametrize("name", ["rect_waveguide", "cube_cavity", "wr42_waveguide", "coax_50ohm"])
def test_fixtures_bit_exact(ctx, name):
    check(ctx, load_fixture_mesh(name, fast=False))  # literal dict-walk restatement of the reference


def test_shuffled_sparse_ids_and_tri_only_edges(ctx):
    rng = np.random.default_rng(5)
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(5, jitter=0.1)
    n = xyz.shape[0]
    ids = rng.permutation(np.arange(1, 40 * n, 7)[:n]).astype(np.int64)  # sparse, shuffled node ids (orientation follows ids)
    # a few triangles that are faces of no tet: their edges are numbered after all tet edges
    extra = np.array([[1, n, n // 2], [2, n - 1, n // 3]], dtype=np.int64)  # meshgen ids are 1-based
    tris2 = np.vstack([tris, extra])
    trp2 = np.concatenate([trp, [77, 77]])
    om = orc.mesh_from_arrays(xyz, ids[tets - 1], tp, ids[tris2 - 1], trp2, node_ids=ids, fast=False)
    assert om.tri_edges.max() >= om.tet_edges.max()  # the tri-only edges exist
    check(ctx, om)
    # the fast (vectorised) oracle agrees too
    om2 = orc.mesh_from_arrays(xyz, ids[tets - 1], tp, ids[tris2 - 1], trp2, node_ids=ids, fast=True)
    check(ctx, om2)


def test_tets_only_empty_and_bad_ids(ctx):
    te, to, re_, ro, edges = cabi.build_edges_device(ctx, np.array([[5, 3, 9, 1]]))
    assert te.tolist() == [[0, 1, 2, 3, 4, 5]] and to.tolist() == [[-1, 1, -1, 1, -1, -1]]
    assert edges.tolist() == [[3, 5], [5, 9], [1, 5], [3, 9], [1, 3], [1, 9]] and re_.shape == (0, 3)
    te, to, re_, ro, edges = cabi.build_edges_device(ctx, np.zeros((0, 4), dtype=np.int64))
    assert te.shape == (0, 6) and edges.shape == (0, 2)
    with pytest.raises(cabi.EfbError):
        cabi.build_edges_device(ctx, np.array([[1, 2, 3, 1 << 33]]))


def test_larger_cube_matches_host_numbering(ctx):
    """120 k tets: device numbering == the C++ host walk (pyedgefem.mesh_from_arrays) == the oracle rules."""
    import edgefem_b200

    pe = edgefem_b200.load_pyedgefem()
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(27, jitter=0.1)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    te, to, re_, ro, edges = cabi.build_edges_device(ctx, tets, tris)  # meshgen conn = node ids 1..n, like mesh_from_arrays
    assert np.array_equal(te, hm.tet_edges_array()) and np.array_equal(to, hm.tet_orient_array())
    assert np.array_equal(re_, hm.tri_edges_array()) and np.array_equal(ro, hm.tri_orient_array())
    assert np.array_equal(edges, hm.edges_array())


def test_device_mesh_from_conn_equals_host_ingest(ctx):
    """The array-only large-mesh ingest (GPU numbering + PEC flags from the boundary tris) gives the same device
    inputs as the Mesh-based host ingest: edge->node table, PEC flags, and the CSR pattern built from them."""
    import edgefem_b200

    pe = edgefem_b200.load_pyedgefem()
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(9, jitter=0.1)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    bc = pe.build_edge_pec(hm, 1)
    dm, info = cabi.device_mesh_from_conn(ctx, xyz, tets, tp, tris)
    assert np.array_equal(info["edge_nodes"], hm.edge_nodes_array())
    flags = cabi.pec_flags_from_tris(info["edges"].shape[0], info["tri_edges"], trp, 1)
    want = np.zeros_like(flags)
    want[np.asarray(bc.dirichlet_edges, dtype=np.int64)] = 1
    assert np.array_equal(flags, want)
    dm_h = cabi.DeviceMesh(ctx, hm.xyz_array(), hm.tet_nodes_array(), hm.tet_edges_array(), hm.tet_orient_array(), hm.tet_phys_array(),
                           hm.edge_nodes_array())
    s1, s2 = cabi.DeviceSystem.from_mesh(dm), cabi.DeviceSystem.from_mesh(dm_h)
    (rp1, ci1), (rp2, ci2) = s1.pattern(), s2.pattern()
    assert np.array_equal(rp1, rp2) and np.array_equal(ci1, ci2)
    for s in (s1, s2):
        s.close()
    dm.close()
    dm_h.close()


def test_device_setup_equals_host_setup(ctx, monkeypatch):
    """Large-mesh set-up on the device (incidence lists by stable radix sort, CSR pattern + position map by a per-row
    kernel) produces the same pattern and, through the same assembly kernel, bit-identical matrix values as the host
    set-up -- whole mesh and a row block."""
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(12, jitter=0.1)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("EDGEFEM_B200_DEVICE_SETUP", mode)
        dm, info = cabi.device_mesh_from_conn(ctx, xyz, tets, tp, tris)
        m = info["edges"].shape[0]
        flags = cabi.pec_flags_from_tris(m, info["tri_edges"], trp, 1)
        mats, keep = cabi.make_materials(len(dm.slot_tags), eps=[2.0 - 0.1j])
        res = []
        for (a, b) in ((0, m), cabi.dist_row_range(m, 1, 3)):
            s = cabi.DeviceSystem.from_mesh(dm) if (a, b) == (0, m) else cabi.DeviceSystem.from_mesh_rows(dm, a, b)
            s.set_dirichlet(flags)
            s.assemble_volume([2 * np.pi * 3e8], mats)
            rp, ci = s.pattern()
            res.append((rp, ci, s.values(0)))
            s.close()
        out[mode] = res
        dm.close()
    for (rp0, ci0, v0), (rp1, ci1, v1) in zip(out["0"], out["1"]):
        assert np.array_equal(rp0, rp1) and np.array_equal(ci0, ci1)
        assert np.array_equal(v0, v1) and np.any(v0 != 0)
