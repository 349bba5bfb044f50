"""CPU tests of the C++ host layer (pyedgefem) against the oracle: mesh ingest, edge numbering
(bit-exact), PEC set, port builders, materials, periodic pairing.  No GPU needed."""
import math
import os

import numpy as np
import pytest

import edgefem_oracle as orc
from conftest import GOLDEN, load_fixture_mesh
from edgefem_b200 import meshgen, load_pyedgefem

pe = load_pyedgefem()


def fixture_arrays(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z


def host_mesh(name):
    z = fixture_arrays(name)
    return pe.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], z["node_ids"].tolist())


@pytest.mark.parametrize("name", ["rect_waveguide", "cube_cavity", "wr42_waveguide", "coax_50ohm"])
def test_edge_numbering_bit_exact(name):
    """build_edges: first-seen numbering over tets then tris, orientation by node id (src/mesh_gmsh.cpp:104-146)."""
    om = load_fixture_mesh(name, fast=False)  # literal dict-walk restatement
    hm = host_mesh(name)
    assert hm.num_edges() == om.num_edges
    assert np.array_equal(hm.tet_edges_array(), om.tet_edges)
    assert np.array_equal(hm.tet_orient_array(), om.tet_orient)
    assert np.array_equal(hm.tri_edges_array(), om.tri_edges)
    assert np.array_equal(hm.tri_orient_array(), om.tri_orient)
    assert np.array_equal(hm.edges_array(), om.edges)


def test_edge_rules_reference_test_edge_indexing():
    """tests/test_edge_indexing.cpp:13-30: key symmetric, edges stored (min,max), orient sign rule."""
    hm = host_mesh("cube_cavity")
    e = hm.edges_array()
    assert np.all(e[:, 0] < e[:, 1])
    t0 = hm.tets[0]
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for k, (a, b) in enumerate(pairs):
        na, nb = t0.conn[a], t0.conn[b]
        assert t0.edge_orient[k] == (1 if na < nb else -1)
        assert tuple(e[t0.edges[k]]) == (min(na, nb), max(na, nb))
    assert orc.make_edge_key(3, 7) == orc.make_edge_key(7, 3)


def test_gmsh_reader_roundtrip(tmp_path, kat):
    z = fixture_arrays("rect_waveguide")
    path = str(tmp_path / "wg.msh")
    meshgen.write_gmsh_v2(path, z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], z["node_ids"])
    hm = pe.load_gmsh(path)
    om = orc.load_gmsh_v2(path)
    c = kat["wr90_counts"]
    assert hm.num_tets() == c["tets"] and hm.num_edges() == c["edges"]
    assert np.array_equal(hm.tet_edges_array(), om.tet_edges)
    assert np.array_equal(hm.edges_array(), om.edges)
    bc = pe.build_edge_pec(hm, 1)
    assert len(bc.dirichlet_edges) == c["pec_edges"]
    assert set(bc.dirichlet_edges) == orc.build_edge_pec(om, 1)
    with pytest.raises(RuntimeError):
        pe.load_gmsh(str(tmp_path / "missing.msh"))


def test_te10_mode_and_port_2d():
    om = load_fixture_mesh("rect_waveguide")
    hm = host_mesh("rect_waveguide")
    pec = orc.build_edge_pec(om, 1)
    bc = pe.build_edge_pec(hm, 1)
    f = 10e9
    mo = orc.solve_te10_mode(0.02286, 0.01016, f)
    mh = pe.solve_te10_mode(pe.RectWaveguidePort(0.02286, 0.01016), f)
    for k in ("fc", "kc", "omega"):
        assert getattr(mh, k) == pytest.approx(getattr(mo, k), rel=1e-15)
    assert complex(mh.Z0) == pytest.approx(complex(mo.Z0), rel=1e-15)
    assert complex(mh.beta) == pytest.approx(complex(mo.beta), rel=1e-15)
    kc_sq = (math.pi / 0.02286) ** 2
    for tag in (2, 3):
        po = orc.build_wave_port_2d(om, tag, mo, pec, kc_sq)
        ph = pe.build_wave_port_2d(hm, tag, mh, set(bc.dirichlet_edges), kc_sq)
        assert list(ph.edges) == po.edges
        assert ph.mode.kc == pytest.approx(po.mode.kc, rel=1e-10)
        wo, wh = np.asarray(po.weights), np.asarray(ph.weights)
        s = np.sign(np.real(np.vdot(wo, wh)))
        assert np.max(np.abs(wh - s * wo)) < 1e-8 * np.max(np.abs(wo))
        # deterministic sign convention of this build: largest |component| positive
        assert np.real(wh[np.argmax(np.abs(wh))]) > 0


def test_port_surface_mass_matches_oracle():
    om = load_fixture_mesh("rect_waveguide")
    hm = host_mesh("rect_waveguide")
    pec = orc.build_edge_pec(om, 1)
    for tag in (2, 3):
        Mo = orc.assemble_port_surface_mass(om, tag, pec)
        rp, ci, va = pe.assemble_port_surface_mass(hm, tag, pec).to_csr()
        assert np.array_equal(rp, Mo.indptr) and np.array_equal(ci, Mo.indices)
        assert np.max(np.abs(va - Mo.data)) <= 1e-15 * np.max(np.abs(Mo.data))
        assert Mo.nnz == 462


def test_triangle_mass_closed_form_vs_quadrature():
    """tests/test_triangle_mass_matrix.cpp:74 (closed form vs quadrature < 1e-12)."""
    rng = np.random.default_rng(3)
    for _ in range(5):
        v = rng.standard_normal((3, 3))
        M = np.array(pe.triangle_whitney_mass_matrix([list(v[0]), list(v[1]), list(v[2])]))
        assert np.max(np.abs(M - orc.triangle_mass_quadrature(v))) < 1e-12
        assert np.max(np.abs(M - orc.triangle_whitney_mass_matrix(v))) < 1e-14


def test_element_matrices_match_oracle():
    rng = np.random.default_rng(4)
    for _ in range(5):
        X = rng.standard_normal((4, 3))
        K = np.array(pe.whitney_curl_curl_matrix([list(x) for x in X]))
        M = np.array(pe.whitney_mass_matrix([list(x) for x in X]))
        assert np.max(np.abs(K - orc.whitney_curl_curl_matrix(X))) <= 1e-13 * np.max(np.abs(K))
        assert np.max(np.abs(M - orc.whitney_mass_matrix(X))) <= 1e-13 * np.max(np.abs(M))
        assert np.allclose(K, K.T) and np.allclose(M, M.T)


def test_lumped_port_matches_oracle():
    xyz, tets, tp, tris, trp = meshgen.rect_waveguide(nx=3, ny=2, nz=4)
    om = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    for mode, si in ((pe.LumpedPortWeightMode.SurfaceIntegral, True), (pe.LumpedPortWeightMode.Projection, False)):
        cfg = pe.LumpedPortConfig()
        cfg.surface_tag = 2
        cfg.z0 = 50.0
        cfg.e_direction = [0.0, 1.0, 0.0]
        cfg.weight_mode = mode
        ph = pe.build_lumped_port(hm, cfg)
        po = orc.build_lumped_port(om, 2, 50.0, (0.0, 1.0, 0.0), surface_integral=si)
        assert list(ph.edges) == po.edges
        assert np.max(np.abs(np.asarray(ph.weights) - po.weights)) < 1e-13
        assert np.linalg.norm(ph.weights) == pytest.approx(math.sqrt(50.0), rel=1e-14)
    cfg.surface_tag = 999
    with pytest.raises(RuntimeError):
        pe.build_lumped_port(hm, cfg)


def test_dispersive_materials_match_oracle_and_limits():
    """include/edgefem/materials/dispersive.hpp + tests/test_dispersive_materials.cpp limits."""
    m = pe.materials
    w = 2 * math.pi * 10e9
    cases = [
        (m.DebyeMaterial(80.1, 4.9, 9.3e-12), orc.DebyeMaterial(80.1, 4.9, 9.3e-12)),
        (m.DrudeMaterial(1.37e16, 4.05e13), orc.DrudeMaterial(1.37e16, 4.05e13)),
    ]
    lh, lo = m.LorentzMaterial(2.0), orc.LorentzMaterial(2.0)
    for p in ((5.0, 2 * math.pi * 10e9, 2 * math.pi * 0.5e9), (1.5, 2 * math.pi * 30e9, 1e9)):
        lh.add_pole(*p)
        lo.add_pole(*p)
    cases.append((lh, lo))
    dh, do = m.DrudeLorentzMaterial(3.5, 2 * math.pi * 5e9, 1e9), orc.DrudeLorentzMaterial(3.5, 2 * math.pi * 5e9, 1e9)
    dh.add_lorentz_pole(0.8, 2 * math.pi * 14e9, 3e9)
    do.add_lorentz_pole(0.8, 2 * math.pi * 14e9, 3e9)
    cases.append((dh, do))
    for h, o in cases:
        for ww in (w, 0.3 * w, 0.0):
            assert complex(h.eval_eps(ww)) == pytest.approx(complex(o.eval_eps(ww)), rel=1e-14)
        assert complex(h.eval_mu(w)) == 1.0
    assert complex(cases[0][0].eval_eps(0.0)).real == pytest.approx(80.1)  # Debye static limit
    assert complex(cases[1][0].eval_eps(0.0)).real == -1e30  # Drude DC convention
    assert complex(cases[0][0].eval_eps(w)).imag < 0  # loss => Im eps < 0
    for bad in (lambda: m.DebyeMaterial(1, 1, 0.0), lambda: m.DrudeMaterial(0.0, 1.0), lambda: m.DrudeMaterial(1.0, -1.0),
                lambda: m.LorentzMaterial().add_pole(1.0, 0.0, 0.0), lambda: m.DrudeLorentzMaterial(1.0, 1.0, -1.0)):
        with pytest.raises(ValueError):
            bad()
    p = pe.MaxwellParams()
    p.eps_r = 2.0
    p.set_eps_r_region(7, 3.0 - 0.1j)
    p.set_eps_model(8, dh)
    assert p.get_eps_r(7) == 3.0 - 0.1j and p.get_eps_r(5) == 2.0
    assert p.get_eps_r_at_freq(8, w) == pytest.approx(complex(do.eval_eps(w)))
    assert p.get_eps_r_at_freq(7, w) == 3.0 - 0.1j


def test_floquet_phase_formulas():
    """tests/test_periodic.cpp:21,41,64,155 -- phase formulas."""
    pbc = pe.PeriodicBC()
    pbc.period_vector = [0.005, 0.0, 0.0]
    pe.set_floquet_phase(pbc, (100.0, 50.0))
    assert complex(pbc.phase_shift) == pytest.approx(complex(math.cos(0.5), math.sin(0.5)))
    k0 = 2 * math.pi * 10e9 / orc.C0
    ph = pe.floquet_phase_from_angle([0.005, 0.0, 0.0], math.radians(30), 0.0, k0)
    assert complex(ph) == pytest.approx(orc.floquet_phase_from_angle((0.005, 0, 0), math.radians(30), 0.0, k0))
    assert abs(complex(pe.floquet_phase_from_angle([0.005, 0.0, 0.0], 0.0, 0.0, k0)) - 1.0) < 1e-15


def test_periodic_pairs_structured():
    g = np.linspace(0, 1, 4)
    xyz, tets, _ = meshgen.box_grid(g, g, g)
    tris, _ = meshgen.boundary_faces(tets)
    tags = np.full(tris.shape[0], 1, dtype=np.int32)
    tags[meshgen.faces_on_plane(xyz, tris, 0, 0.0)] = 5
    tags[meshgen.faces_on_plane(xyz, tris, 0, 1.0)] = 6
    tp = np.full(tets.shape[0], 100, dtype=np.int32)
    om = orc.mesh_from_arrays(xyz, tets, tp, tris, tags)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, tags)
    po = orc.build_periodic_pairs(om, 5, 6, (1.0, 0.0, 0.0))
    ph = pe.build_periodic_pairs(hm, 5, 6, [1.0, 0.0, 0.0])
    assert pe.validate_periodic_bc(hm, ph)
    assert [(p.master_edge, p.slave_edge, p.master_orient, p.slave_orient) for p in ph.pairs] == \
        [(p.master_edge, p.slave_edge, p.master_orient, p.slave_orient) for p in po.pairs]
    assert pe.count_surface_edges(hm, 5) == len(po.pairs)
    with pytest.raises(RuntimeError):
        pe.build_periodic_pairs(hm, 5, 6, [0.5, 0.0, 0.0])
    with pytest.raises(RuntimeError):
        pe.build_periodic_pairs(hm, 77, 6, [1.0, 0.0, 0.0])


def test_compute_calls_fail_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device every compute entry point raises."""
    if pe.b200_device_count() > 0:
        pytest.skip("a GPU is present")
    hm = host_mesh("cube_cavity")
    bc = pe.build_edge_pec(hm, 1)
    p = pe.MaxwellParams()
    p.omega = 1e9
    with pytest.raises(RuntimeError, match="CUDA"):
        pe.assemble_maxwell(hm, p, bc)
    A = pe.SpMatC.from_csr(2, [0, 1, 2], [0, 1], [1.0, 1.0])
    with pytest.raises(RuntimeError, match="CUDA"):
        pe.solve_linear(A, np.ones(2, dtype=complex))
