"""GPU parity tests through the public API (pyedgefem: C++ host -> C-ABI -> CUDA) vs the oracle."""
import math
import os

import numpy as np
import pytest
import scipy.sparse as sp

import edgefem_oracle as orc
import helpers as H
from conftest import GOLDEN, load_fixture_mesh
from edgefem_b200 import meshgen, load_pyedgefem

pytestmark = pytest.mark.gpu
pe = load_pyedgefem()


def both_meshes(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    hm = pe.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], z["node_ids"].tolist())
    return hm, load_fixture_mesh(name)


def to_host_port(po: orc.WavePort):
    ph = pe.WavePort()
    ph.surface_tag = po.surface_tag
    ph.edges = list(po.edges)
    ph.weights = np.asarray(po.weights, dtype=complex)
    md = pe.PortMode()
    md.kc, md.fc, md.omega = po.mode.kc, po.mode.fc, po.mode.omega
    md.Z0, md.beta, md.eps, md.mu = po.mode.Z0, po.mode.beta, po.mode.eps, po.mode.mu
    ph.mode = md
    return ph


def csr(A):
    rp, ci, va = A.to_csr()
    return sp.csr_matrix((va, ci, rp), shape=A.shape)


def make_params(po: orc.MaxwellParams):
    p = pe.MaxwellParams()
    p.omega, p.eps_r, p.mu_r = po.omega, po.eps_r, po.mu_r
    p.eps_r_regions, p.mu_r_regions = dict(po.eps_r_regions), dict(po.mu_r_regions)
    p.use_abc, p.abc_surface_tags = po.use_abc, set(po.abc_surface_tags)
    p.use_port_abc, p.port_abc_scale = po.use_port_abc, po.port_abc_scale
    p.port_abc_type = [pe.PortABCType.None_, pe.PortABCType.Beta, pe.PortABCType.BetaNorm, pe.PortABCType.ImpedanceMatch,
                       pe.PortABCType.ModalAdmittance][po.port_abc_type]
    p.pml_sigma, p.pml_regions = po.pml_sigma, set(po.pml_regions)
    p.enforce_pml_heuristics = po.enforce_pml_heuristics
    return p


def check_assembly(hm, om, ph_params, po_params, bc, pec, ports_h=(), ports_o=(), active=-1, tol=1e-12):
    asm = pe.assemble_maxwell(hm, ph_params, bc, list(ports_h), active)
    A_o, b_o = orc.assemble_maxwell(om, po_params, pec, list(ports_o), active)
    A_h = csr(asm.A)
    assert np.array_equal(A_h.indptr, A_o.indptr) and np.array_equal(A_h.indices, A_o.indices), "CSR pattern differs"
    err = H.row_rel_err(A_h, A_o)  # 1e-12 relative on the row scale (fp64 tolerance of north_star)
    bh = asm.b.to_numpy()
    berr = np.max(np.abs(bh - b_o)) / max(1e-300, np.max(np.abs(b_o))) if np.max(np.abs(b_o)) > 0 else np.max(np.abs(bh))
    assert err < tol, err
    assert berr < 1e-13, berr
    return A_h, A_o


def test_assemble_maxwell_vacuum_and_lossy():
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    for eps in (1.0, 2.2 - 0.05j):
        po = orc.MaxwellParams(omega=2 * math.pi * 10e9, eps_r=eps, mu_r=1.0 if eps == 1.0 else 1.1 - 0.01j)
        A_h, A_o = check_assembly(hm, om, make_params(po), po, bc, pec)
        # tests/test_maxwell.cpp:16,104: lossless real-symmetric => Hermitian; lossy not
        herm = abs(A_h - A_h.conj().T).max()
        assert (herm < 1e-9) == (eps == 1.0)
        assert abs(A_h - A_h.T).max() < 1e-9 * abs(A_h).max()  # always complex symmetric


def test_assemble_maxwell_ports_abc_regions():
    """Dense port block (new pattern entries), RHS 2w/sqrt(Z0), ABC diag, port ABC, region materials."""
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    f = 10e9
    ports_o = orc.wr90_ports(om, pec, f)
    for p_ in ports_o:
        p_.weights = p_.weights * (1.3 + 0j)
    ports_h = [to_host_port(p_) for p_ in ports_o]
    for abc_type in (orc.PORT_ABC_BETA, orc.PORT_ABC_BETA_NORM, orc.PORT_ABC_IMPEDANCE_MATCH, orc.PORT_ABC_MODAL_ADMITTANCE):
        po = orc.MaxwellParams(omega=2 * math.pi * f, use_abc=True, abc_surface_tags={2}, use_port_abc=True, port_abc_type=abc_type,
                               eps_r_regions={100: 1.5 - 0.02j})
        A_h, A_o = check_assembly(hm, om, make_params(po), po, bc, pec, ports_h, ports_o, active=1)
    assert A_o.nnz > 85113  # the dense block added entries
    # tests/test_maxwell.cpp:38-41: b(edge) = 2 w / sqrt(Z0)
    asm = pe.assemble_maxwell(hm, make_params(po), bc, ports_h, 0)
    b = asm.b.to_numpy()
    k = int(np.argmax(np.abs(ports_o[0].weights)))
    assert b[ports_o[0].edges[k]] == pytest.approx(2.0 * ports_o[0].weights[k] / np.sqrt(ports_o[0].mode.Z0), rel=1e-14)


def test_assemble_maxwell_dispersive_and_pml():
    g = np.linspace(0.0, 1.0, 6)
    xyz, tets, cells = meshgen.box_grid(g, g * 0.8, g * 1.2)
    rng = np.random.default_rng(7)
    interior = np.all((xyz > 1e-9) & (xyz < np.array([1.0, 0.8, 1.2]) - 1e-9), axis=1)
    xyz[interior] += (rng.random((int(interior.sum()), 3)) - 0.5) * 0.06
    tp = (100 + np.minimum(cells[:, 2] // 2, 2)).astype(np.int32)  # three material layers 100,101,102
    tris, _ = meshgen.boundary_faces(tets)
    trp = np.ones(tris.shape[0], dtype=np.int32)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    om = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
    tags = [100, 101, 102]
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    w = 2 * math.pi * 1.0e9
    m = pe.materials
    # all four device model kinds (EFB_MODEL_DEBYE / DRUDE_LORENTZ / LORENTZ with two poles / DRUDE), one at a time ...
    models_h = [m.DebyeMaterial(8.0, 3.0, 2e-10), m.DrudeLorentzMaterial(3.5, 2 * math.pi * 0.4e9, 5e8), m.LorentzMaterial(2.2),
                m.DrudeMaterial(2 * math.pi * 0.6e9, 3e8)]
    models_o = [orc.DebyeMaterial(8.0, 3.0, 2e-10), orc.DrudeLorentzMaterial(3.5, 2 * math.pi * 0.4e9, 5e8), orc.LorentzMaterial(2.2),
                orc.DrudeMaterial(2 * math.pi * 0.6e9, 3e8)]
    models_h[1].add_lorentz_pole(0.7, 2 * math.pi * 2.0e9, 4e8)
    models_o[1].add_lorentz_pole(0.7, 2 * math.pi * 2.0e9, 4e8)
    for mdl in (models_h[2], models_o[2]):
        mdl.add_pole(1.3, 2 * math.pi * 1.7e9, 2.5e8)
        mdl.add_pole(0.4, 2 * math.pi * 3.1e9, 6e8)
    for k in range(4):
        tag = tags[k % 3]
        assert abs(models_h[k].eval_eps(w) - models_o[k].eval_eps(w)) <= 1e-14 * abs(models_o[k].eval_eps(w))
        po = orc.MaxwellParams(omega=w, eps_models={tag: models_o[k]})
        ph = make_params(po)
        ph.set_eps_model(tag, models_h[k])
        check_assembly(hm, om, ph, po, bc, pec)
    # ... and three different kinds in the three layers of one assembly (the per-slot table of the kernel)
    po = orc.MaxwellParams(omega=w, eps_models={tags[0]: models_o[2], tags[1]: models_o[3], tags[2]: models_o[0]})
    ph = make_params(po)
    for tag, mdl in zip(tags, (models_h[2], models_h[3], models_h[0])):
        ph.set_eps_model(tag, mdl)
    check_assembly(hm, om, ph, po, bc, pec)
    # uniform PML region + tensor PML region
    po = orc.MaxwellParams(omega=w, pml_sigma=2.0e9, pml_regions={tags[0]})
    check_assembly(hm, om, make_params(po), po, bc, pec)
    spec_o = orc.PMLRegionSpec(sigma_max=(3e9, 0.0, 1e9), thickness=(0.3, 0.0, 0.2), grading_order=2.0)
    po = orc.MaxwellParams(omega=w, pml_tensor_regions={tags[-1]: spec_o})
    ph = make_params(po)
    sp_h = pe.PMLRegionSpec()
    sp_h.sigma_max, sp_h.thickness, sp_h.grading_order = [3e9, 0.0, 1e9], [0.3, 0.0, 0.2], 2.0
    ph.pml_tensor_regions = {tags[-1]: sp_h}
    check_assembly(hm, om, ph, po, bc, pec)
    asm = pe.assemble_maxwell(hm, ph, bc)
    d_o = orc.pml_diagnostics(po)[0]
    assert np.allclose(asm.diagnostics[0].reflection_est, d_o[3], rtol=1e-14)


def test_solve_linear_contract():
    hm, om = both_meshes("rect_waveguide")
    pec = orc.build_edge_pec(om, 1)
    ports = orc.wr90_ports(om, pec, 10e9)
    A, pm, pv, betas = orc.eigenmode_system(om, orc.MaxwellParams(omega=2 * math.pi * 10e9), pec, ports)
    b = 2j * betas[0] * (pm[0] @ pv[0])
    Ah = pe.SpMatC.from_csr(A.shape[0], A.indptr, A.indices, A.data)
    x_ref, _ = orc.solve_direct(A, b)
    opts = pe.SolveOptions()
    res = pe.solve_linear(Ah, b, opts)
    assert res.converged and res.residual <= 1e-10 * 1.001 and res.iters > 0
    assert res.method.startswith("B200:COCG")  # complex symmetric => COCG; no gradient => Jacobi
    x = res.x.to_numpy()
    assert np.linalg.norm(x - x_ref) <= 1e-7 * np.linalg.norm(x_ref)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) == pytest.approx(res.residual, rel=1e-3)
    # non-symmetric (diagonally dominant) matrix => BiCGSTAB branch
    rng = np.random.default_rng(5)
    n = A.shape[0]
    R = sp.random(n, n, density=4.0 / n, random_state=rng, format="csr")
    R.data = R.data + 1j * rng.standard_normal(R.nnz)
    B = (R + sp.diags(np.full(n, 12.0 + 3.0j))).tocsr()
    B.sort_indices()
    Bh = pe.SpMatC.from_csr(n, B.indptr, B.indices, B.data.astype(complex))
    r2 = pe.solve_linear(Bh, b)
    assert r2.method.startswith("B200:BiCGSTAB") and r2.converged
    assert np.linalg.norm(B @ r2.x.to_numpy() - b) / np.linalg.norm(b) < 2e-10
    # max_iterations exhausted => converged False + message, no exception
    opts.max_iterations = 5
    opts.auto_fallback = False
    r3 = pe.solve_linear(Ah, b, opts)
    assert not r3.converged and r3.error_message and r3.iters <= 5 + 1
    # zero rhs
    r4 = pe.solve_linear(Ah, np.zeros_like(b))
    assert r4.converged and np.all(r4.x.to_numpy() == 0)


def test_calculate_sparams_eigenmode_kat_table(kat):
    """docs/validation.md:22-27 through the public API with this build's own port builder."""
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    kc_sq = (math.pi / 0.02286) ** 2
    dims = pe.RectWaveguidePort(0.02286, 0.01016)
    for f_ghz, s11, s21, ph21 in kat["wr90_table"]["rows"]:
        f = f_ghz * 1e9
        ports = [pe.build_wave_port_2d(hm, tag, pe.solve_te10_mode(dims, f), set(bc.dirichlet_edges), kc_sq) for tag in (2, 3)]
        p = pe.MaxwellParams()
        p.omega = 2 * math.pi * f
        S = pe.calculate_sparams_eigenmode(hm, p, bc, ports)
        assert abs(abs(S[0, 0]) - s11) < 1e-3
        assert abs(abs(S[1, 0]) - s21) < 1e-4
        # the port eigenvector sign is implementation-defined: S21 matches the table up to a sign
        d = abs(((np.angle(S[1, 0], deg=True) - ph21 + 180) % 360) - 180)
        assert min(d, abs(d - 180)) < 0.06
        # against the oracle with the SAME ports: 1e-6
        S_o = orc.calculate_sparams_eigenmode(om, orc.MaxwellParams(omega=p.omega), pec, [orc.WavePort(
            surface_tag=q.surface_tag, mode=orc.PortMode(kc=q.mode.kc), edges=list(q.edges), weights=np.asarray(q.weights)) for q in ports])
        assert np.max(np.abs(S - S_o)) <= 1e-6
        # enforced thresholds tests/benchmark_wr90.cpp:91
        assert abs(S[0, 0]) < 0.15 and abs(S[1, 0]) > 0.90 and abs(S[0, 0]) ** 2 + abs(S[1, 0]) ** 2 <= 1.05


def test_per_frequency_calls_reuse_the_device_system():
    """calculate_sparams_eigenmode in a frequency loop (python/edgefem/designs/waveguide.py:394-433) keeps the device system
    and the port surface matrices of the previous call: a call must not see what the one before left behind."""
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    dims = pe.RectWaveguidePort(0.02286, 0.01016)
    ports = [pe.build_wave_port_2d(hm, tag, pe.solve_te10_mode(dims, 10e9), set(bc.dirichlet_edges), (math.pi / 0.02286) ** 2) for tag in (2, 3)]

    def call(f, scale=1.0):
        p = pe.MaxwellParams()
        p.omega = 2 * math.pi * f
        p.port_abc_scale = scale
        return np.array(pe.calculate_sparams_eigenmode(hm, p, bc, ports))

    chain = [call(9e9), call(11.5e9), call(9e9, 0.5), call(9e9)]
    fresh = []
    for f, sc in ((9e9, 1.0), (11.5e9, 1.0), (9e9, 0.5), (9e9, 1.0)):
        pe.b200_clear_cache()
        fresh.append(call(f, sc))
    # repeats were measured bit-identical once the preconditioner's nodal diagonal and the port right-hand side stopped using
    # fp64 atomics; the bound stays at the solver tolerance
    for a, b in zip(chain, fresh):
        assert np.max(np.abs(a - b)) <= 1e-9
    assert np.max(np.abs(chain[0] - chain[3])) <= 1e-9 and np.max(np.abs(chain[0] - chain[2])) > 1e-3
    # another Dirichlet set on the same mesh takes another system
    bc2 = pe.build_edge_pec(hm, 1)
    bc2.merge(pe.build_edge_pec(hm, 3))
    ports2 = [pe.build_wave_port_2d(hm, 2, pe.solve_te10_mode(dims, 10e9), set(bc2.dirichlet_edges), (math.pi / 0.02286) ** 2)]
    p = pe.MaxwellParams()
    p.omega = 2 * math.pi * 10e9
    S_short = np.array(pe.calculate_sparams_eigenmode(hm, p, bc2, ports2))
    assert abs(abs(S_short[0, 0]) - 1.0) < 2e-3  # shorted guide: total reflection (measured 0.9996)
    assert np.max(np.abs(call(9e9) - chain[0])) <= 1e-9


def test_alpha_sweep_kat(kat):
    """docs/validation.md:51-57 via port_abc_scale."""
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    f = 10e9
    dims = pe.RectWaveguidePort(0.02286, 0.01016)
    ports = [pe.build_wave_port_2d(hm, tag, pe.solve_te10_mode(dims, f), set(bc.dirichlet_edges), (math.pi / 0.02286) ** 2) for tag in (2, 3)]
    for alpha, s11, s21 in kat["alpha_sweep_10ghz"]["rows"]:
        p = pe.MaxwellParams()
        p.omega = 2 * math.pi * f
        p.port_abc_scale = alpha
        S = pe.calculate_sparams_eigenmode(hm, p, bc, ports)
        assert abs(abs(S[0, 0]) - s11) < 1e-3 and abs(abs(S[1, 0]) - s21) < 1e-3


def test_eigenmode_sweep_batch_and_evanescent():
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    ports_o = orc.wr90_ports(om, pec, 10e9)
    ports_h = [to_host_port(q) for q in ports_o]
    freqs = list(np.linspace(8e9, 12e9, 5)) + [5e9]  # last one is below cutoff (6.56 GHz)
    S, st = pe.calculate_sparams_eigenmode_sweep(hm, pe.MaxwellParams(), bc, ports_h, freqs)
    assert len(S) == 6 and st.kernel_launches > 0 and st.device_ms > 0
    for fi, f in enumerate(freqs[:-1]):
        S_o = orc.wr90_sparams(om, pec, f, ports_o)
        assert np.max(np.abs(S[fi] - S_o)) <= 1e-6
    assert np.all(np.isnan(S[-1]))
    assert all(st.converged[: 2 * 5])


def test_calculate_sparams_lumped_and_normalize():
    """calculate_sparams + normalize_port_weights with a lumped port, ABC and a lossy substrate region."""
    xyz, tets, tp, tris, trp = meshgen.rect_waveguide(a=0.02, b=0.01, length=0.03, nx=6, ny=3, nz=9)
    trp = trp.copy()
    trp[trp == 3] = 50  # far end radiates (ABC), near end is the lumped port (tag 2)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    om = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    po = orc.MaxwellParams(omega=2 * math.pi * 9e9, use_abc=True, abc_surface_tags={50}, eps_r_regions={100: 2.2 * (1 - 0.02j)})
    ph = make_params(po)
    cfg = pe.LumpedPortConfig()
    cfg.surface_tag, cfg.z0, cfg.e_direction = 2, 50.0, [0.0, 1.0, 0.0]
    port_h = pe.build_lumped_port(hm, cfg)
    port_o = orc.build_lumped_port(om, 2, 50.0, (0.0, 1.0, 0.0))
    ports_h = pe.normalize_port_weights(hm, ph, bc, [port_h])
    ports_o = [port_o]
    orc.normalize_port_weights(om, po, pec, ports_o)
    assert np.max(np.abs(np.asarray(ports_h[0].weights) - ports_o[0].weights)) <= 1e-7 * np.max(np.abs(ports_o[0].weights))
    S_h = pe.calculate_sparams(hm, ph, bc, ports_h)
    S_o = orc.calculate_sparams(om, po, pec, ports_o)
    assert np.max(np.abs(S_h - S_o)) <= 1e-6 * max(1.0, np.max(np.abs(S_o)))
    assert abs(S_h[0, 0]) <= 1.0 + 1e-9


def test_km_and_frequency_sweep():
    xyz, tets, tp, tris, trp = meshgen.rect_waveguide(a=0.02, b=0.01, length=0.03, nx=5, ny=3, nz=8)
    trp = trp.copy()
    trp[trp == 3] = 50
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    om = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    po = orc.MaxwellParams(use_abc=True, abc_surface_tags={50}, eps_r_regions={100: 2.2 - 0.03j}, mu_r=1.05)
    ph = make_params(po)
    km = pe.assemble_maxwell_km(hm, ph, bc)
    K_o, M_o = orc.assemble_maxwell_km(om, po, pec)
    K_h, M_h = csr(km.K), csr(km.M)
    assert np.array_equal(K_h.indices, K_o.indices) and np.array_equal(K_h.indptr, K_o.indptr)
    assert abs(K_h - K_o).max() <= 1e-12 * abs(K_o).max() and abs(M_h - M_o).max() <= 1e-12 * abs(M_o).max()
    w = 2 * math.pi * 9e9
    A_c = csr(km.combine(w))
    assert abs(A_c - (K_o - (w / orc.C0) ** 2 * M_o)).max() <= 1e-12 * abs(K_o).max()
    cfg = pe.LumpedPortConfig()
    cfg.surface_tag, cfg.z0, cfg.e_direction = 2, 50.0, [0.0, 1.0, 0.0]
    ports_h = [pe.build_lumped_port(hm, cfg)]
    ports_o = [orc.build_lumped_port(om, 2, 50.0, (0.0, 1.0, 0.0))]
    freqs = [8e9, 9e9, 10e9]
    sw = pe.frequency_sweep(hm, ph, bc, ports_h, freqs)          # fast K/M path
    S_o = orc.frequency_sweep(om, po, pec, ports_o, freqs)
    assert list(sw.frequencies) == freqs
    for fi in range(3):
        assert np.max(np.abs(sw.S_matrices[fi] - S_o[fi])) <= 1e-6
    # slow path: a dispersive model forces per-frequency assembly (src/sweep.cpp:330-346)
    mo = orc.DebyeMaterial(4.0, 2.0, 1e-11)
    po.eps_models = {100: mo}
    ph.set_eps_model(100, pe.materials.DebyeMaterial(4.0, 2.0, 1e-11))
    sw2 = pe.frequency_sweep(hm, ph, bc, ports_h, freqs)
    S_o2 = orc.frequency_sweep(om, po, pec, ports_o, freqs)
    for fi in range(3):
        assert np.max(np.abs(sw2.S_matrices[fi] - S_o2[fi])) <= 1e-6
    assert len(pe.frequency_sweep(hm, ph, bc, [], freqs).S_matrices) == 0  # src/sweep.cpp:183-185


def test_calculate_sparams_modal_line_integral_ports():
    """SURVEY 8f-f1 feeding the path: ports built by extract_surface_mesh + solve_port_eigens (nodal TE mode) +
    build_wave_port, then the dense-port-block path calculate_sparams on the device vs the oracle with ITS OWN
    restatement of the same builders."""
    hm, om = both_meshes("rect_waveguide")
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    omega = 2 * math.pi * 10e9
    ports_h, ports_o = [], []
    for tag in (2, 3):
        hs, os_ = pe.extract_surface_mesh(hm, tag), orc.extract_surface_mesh(om, tag)
        hmode = pe.solve_port_eigens(hs.mesh, 1, omega, 1.0, 1.0, pe.ModePolarization.TE)[0]
        omode, ofld = orc.solve_port_eigens(os_, 1, omega, 1.0, 1.0)[0]
        ports_h.append(pe.build_wave_port(hm, hs, hmode))
        ports_o.append(orc.build_wave_port(om, os_, omode, ofld))
        wh, wo = np.asarray(ports_h[-1].weights), ports_o[-1].weights
        assert list(ports_h[-1].edges) == ports_o[-1].edges
        assert np.max(np.abs(wh - wo)) <= 1e-7 * np.max(np.abs(wo))
    po = orc.MaxwellParams(omega=omega)
    ph = make_params(po)
    S_h = pe.calculate_sparams(hm, ph, bc, ports_h)
    S_o = orc.calculate_sparams(om, po, pec, ports_o)
    assert np.max(np.abs(S_h - S_o)) <= 1e-6 * max(1.0, np.max(np.abs(S_o)))
