"""GPU parity of the named workloads of BASELINE.json that are not the WR-90 sweep, driven through the public API
exactly as the reference's design classes drive it, against the oracle (SuperLU) on the same synthetic meshes:

 C3  probe-fed patch antenna on FR-4, lumped port + ABC: python/edgefem/designs/stacked_patch.py:469-561 as run by
     examples/run_patch_fullwave.py:19-32 (11 points 2.2-2.7 GHz, use_direct=True, port rebuilt and re-normalised at
     every frequency).  The gmsh model is replaced by meshgen.patch_antenna (same dimensions and tags).
 C4  periodic unit cell, Bloch phase along x, modal top port from solve_port_eigens + build_wave_port, port ABC
     (python/edgefem/designs/unit_cell.py:640-770 as run by examples/unit_cell_demo.py:19-33: 5 points 8-12 GHz),
     with the Drude-Lorentz substrate BASELINE config 4 names (params.set_eps_model), at normal incidence and at
     theta = 30 deg (non-real Bloch phase: the system is not symmetric and takes the BiCGSTAB path).
Parity is "unpinned upstream" for both (the reference holds no golden values for these flows)."""
import math

import numpy as np
import pytest

import edgefem_oracle as orc
from edgefem_b200 import meshgen, load_pyedgefem

pytestmark = pytest.mark.gpu
pe = load_pyedgefem()


def patch_meshes(**kw):
    xyz, tets, tp, tris, trp, info = meshgen.patch_antenna(**kw)
    return pe.mesh_from_arrays(xyz, tets, tp, tris, trp), orc.mesh_from_arrays(xyz, tets, tp, tris, trp), info


def test_c3_patch_antenna_s11_sweep():
    hm, om, info = patch_meshes()
    assert hm.num_tets() > 10000
    bc = pe.BC()
    pec = set()
    for tag in (1, 2, 3, 4, 10):  # ground, cavity walls, cavity bottom, probe, patch (stacked_patch.py:476-490)
        assert pe.has_surface_tag(hm, tag)
        bc.merge(pe.build_edge_pec(hm, tag))
        pec |= orc.build_edge_pec(om, tag)
    assert set(bc.dirichlet_edges) == pec
    eps_sub = complex(4.4, -4.4 * 0.02)
    worst = 0.0
    s11 = []
    for f in np.linspace(2.2e9, 2.7e9, 11):
        ph = pe.MaxwellParams()
        ph.omega = 2 * math.pi * f
        ph.use_abc = True
        ph.abc_surface_tags = {50}
        ph.set_eps_r_region(110, eps_sub)
        po = orc.MaxwellParams(omega=2 * math.pi * f, use_abc=True, abc_surface_tags={50}, eps_r_regions={110: eps_sub})
        cfg = pe.LumpedPortConfig()
        cfg.surface_tag, cfg.z0, cfg.e_direction = 5, 50.0, [1.0, 0.0, 0.0]
        opts = pe.SolveOptions()
        opts.use_direct = True
        ports_h = pe.normalize_port_weights(hm, ph, bc, [pe.build_lumped_port(hm, cfg)], opts)
        ports_o = [orc.build_lumped_port(om, 5, 50.0, (1.0, 0.0, 0.0))]
        orc.normalize_port_weights(om, po, pec, ports_o)
        wo = ports_o[0].weights
        assert np.max(np.abs(np.asarray(ports_h[0].weights) - wo)) <= 1e-6 * np.max(np.abs(wo))
        S_h = pe.calculate_sparams(hm, ph, bc, ports_h, opts)
        S_o = orc.calculate_sparams(om, po, pec, ports_o)
        assert np.all(np.isfinite(S_h)), "solver did not converge at %.3f GHz" % (f / 1e9)
        worst = max(worst, float(np.max(np.abs(S_h - S_o)) / max(1.0, np.max(np.abs(S_o)))))
        s11.append(S_h[0, 0])
    print("C3 patch: %d tets, %d edges, max |S11 - S11_oracle| (relative to max(1,|S|)) over 11 points = %.2e" % (hm.num_tets(), hm.num_edges(), worst))
    assert worst <= 1e-6


def unit_cell_setup(nx=8, ny=8, nz_sub=2, nz_air=8):
    # 5 mm cell, 0.5 mm substrate, air = lambda/2 at 10 GHz (unit_cell.py:109), 4 x 4 mm patch faces tagged 2
    h_air = 299792458.0 / 10e9 / 2
    xyz, tets, tp, tris, trp = meshgen.unit_cell(px=5e-3, py=5e-3, h_sub=0.5e-3, h_air=h_air, nx=nx, ny=ny, nz_sub=nz_sub, nz_air=nz_air,
                                                patch=(4e-3, 4e-3))
    return pe.mesh_from_arrays(xyz, tets, tp, tris, trp), orc.mesh_from_arrays(xyz, tets, tp, tris, trp)


@pytest.mark.parametrize("theta", [0.0, 30.0])
def test_c4_unit_cell_bloch_drude_lorentz(theta):
    hm, om = unit_cell_setup()
    assert hm.num_tets() == 8 * 8 * 10 * 6
    bc = pe.build_edge_pec(hm, 1)
    pec = orc.build_edge_pec(om, 1)
    # builder-chosen substrate dispersion (recorded here): eps_inf 3.5, plasma 2 pi 4 GHz, gamma_d 2 pi 0.5 GHz,
    # one Lorentz pole (delta_eps 0.8, 2 pi 15 GHz, gamma 2 pi 1 GHz)
    dl = dict(eps_inf=3.5, omega_p=2 * math.pi * 4e9, gamma_d=2 * math.pi * 0.5e9, poles=[(0.8, 2 * math.pi * 15e9, 2 * math.pi * 1e9)])
    worst = 0.0
    for f in np.linspace(8e9, 12e9, 5):
        omega = 2 * math.pi * f
        k0 = omega / orc.C0
        kx = k0 * math.sin(math.radians(theta))
        pbc_h = pe.build_periodic_pairs(hm, 5, 6, [5e-3, 0.0, 0.0])
        pe.set_floquet_phase(pbc_h, [kx, 0.0])
        pbc_o = orc.build_periodic_pairs(om, 5, 6, (5e-3, 0.0, 0.0))
        orc.set_floquet_phase(pbc_o, kx, 0.0)
        ph = pe.MaxwellParams()
        ph.omega = omega
        ph.eps_r_regions = {100: complex(3.5, 0.0), 101: complex(1.0, 0.0)}
        ph.use_port_abc = True
        mdl_h = pe.materials.DrudeLorentzMaterial(dl["eps_inf"], dl["omega_p"], dl["gamma_d"])
        for de, w0, g in dl["poles"]:
            mdl_h.add_lorentz_pole(de, w0, g)
        ph.set_eps_model(100, mdl_h)
        po = orc.MaxwellParams(omega=omega, eps_r_regions={100: complex(3.5, 0.0), 101: complex(1.0, 0.0)}, use_port_abc=True)
        mdl_o = orc.DrudeLorentzMaterial(dl["eps_inf"], dl["omega_p"], dl["gamma_d"])
        for de, w0, g in dl["poles"]:
            mdl_o.add_lorentz_pole(de, w0, g)
        po.eps_models = {100: mdl_o}
        hs, os_ = pe.extract_surface_mesh(hm, 4), orc.extract_surface_mesh(om, 4)
        hmode = pe.solve_port_eigens(hs.mesh, 1, omega, 1.0, 1.0, pe.ModePolarization.TE)[0]
        omode, ofld = orc.solve_port_eigens(os_, 1, omega, 1.0, 1.0)[0]
        port_h = pe.build_wave_port(hm, hs, hmode)
        port_o = orc.build_wave_port(om, os_, omode, ofld)
        # degenerate eigen-subspaces of the square cell make the mode solver-dependent: both sides get ONE port
        port_h.weights = np.asarray(port_o.weights, dtype=complex)
        assert list(port_h.edges) == port_o.edges
        S_h = pe.calculate_sparams_periodic(hm, ph, bc, pbc_h, [port_h])
        S_o = orc.calculate_sparams_periodic(om, po, pec, pbc_o, [port_o])
        assert np.all(np.isfinite(S_h)), "solver did not converge at %.1f GHz, theta %.0f" % (f / 1e9, theta)
        worst = max(worst, float(np.max(np.abs(S_h - S_o)) / max(1.0, np.max(np.abs(S_o)))))
    print("C4 unit cell theta=%.0f: %d tets, %d edges, max |R - R_oracle| over 5 points = %.2e" % (theta, hm.num_tets(), hm.num_edges(), worst))
    assert worst <= 1e-6
