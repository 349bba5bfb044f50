"""Periodic (Bloch) elimination: oracle dense (literal reference procedure) vs sparse T A T^H on CPU,
and the device path vs the oracle on the GPU.  No reference test pins this path."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

import edgefem_oracle as orc
from edgefem_b200 import meshgen, load_pyedgefem

pe = load_pyedgefem()


def cell():
    xyz, tets, tp, tris, trp = meshgen.unit_cell(nx=3, ny=3, nz_sub=1, nz_air=3, patch=(0.003, 0.003))
    om = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
    hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
    pec = orc.build_edge_pec(om, 1) | orc.build_edge_pec(om, 2)
    bc = pe.build_edge_pec(hm, 1)
    bc.merge(pe.build_edge_pec(hm, 2))
    return om, hm, pec, bc


def params(theta_deg):
    f = 10e9
    w = 2 * math.pi * f
    dl_o = orc.DrudeLorentzMaterial(3.5, 2 * math.pi * 2e9, 1e9)
    dl_o.add_lorentz_pole(0.6, 2 * math.pi * 18e9, 2e9)
    po = orc.MaxwellParams(omega=w, eps_models={100: dl_o}, use_port_abc=True)
    ph = pe.MaxwellParams()
    ph.omega = w
    dl_h = pe.materials.DrudeLorentzMaterial(3.5, 2 * math.pi * 2e9, 1e9)
    dl_h.add_lorentz_pole(0.6, 2 * math.pi * 18e9, 2e9)
    ph.set_eps_model(100, dl_h)
    ph.use_port_abc = True
    k0 = w / orc.C0
    return po, ph, k0


def test_oracle_sparse_equals_dense_reference_procedure():
    om, hm, pec, bc = cell()
    po, ph, k0 = params(30)
    pbc = orc.build_periodic_pairs(om, 5, 6, (0.005, 0.0, 0.0))
    port = orc.build_lumped_port(om, 4, 376.73, (1.0, 0.0, 0.0))
    port.mode.kc = 0.0
    for theta in (0.0, 30.0):
        pbc.phase_shift = orc.floquet_phase_from_angle((0.005, 0, 0), math.radians(theta), 0.0, k0)
        Ad, bd = orc.assemble_maxwell_periodic_dense(om, po, pec, pbc, [port], 0)
        As, bs = orc.assemble_maxwell_periodic(om, po, pec, pbc, [port], 0)
        assert abs(Ad - As).max() <= 1e-12 * abs(Ad).max()
        assert np.max(np.abs(bd - bs)) <= 1e-13 * np.max(np.abs(bd))
        Sd = orc.calculate_sparams_periodic(om, po, pec, pbc, [port], dense=True)
        Ss = orc.calculate_sparams_periodic(om, po, pec, pbc, [port], dense=False)
        assert np.max(np.abs(Sd - Ss)) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("theta", [0.0, 30.0])
def test_periodic_device_vs_oracle(theta):
    om, hm, pec, bc = cell()
    po, ph, k0 = params(theta)
    pbc_o = orc.build_periodic_pairs(om, 5, 6, (0.005, 0.0, 0.0))
    pbc_h = pe.build_periodic_pairs(hm, 5, 6, [0.005, 0.0, 0.0])
    phase = orc.floquet_phase_from_angle((0.005, 0, 0), math.radians(theta), 0.0, k0)
    pbc_o.phase_shift = phase
    pbc_h.phase_shift = complex(pe.floquet_phase_from_angle([0.005, 0.0, 0.0], math.radians(theta), 0.0, k0))
    port_o = orc.build_lumped_port(om, 4, 376.73, (1.0, 0.0, 0.0))
    cfg = pe.LumpedPortConfig()
    cfg.surface_tag, cfg.z0, cfg.e_direction = 4, 376.73, [1.0, 0.0, 0.0]
    port_h = pe.build_lumped_port(hm, cfg)
    asm = pe.assemble_maxwell_periodic(hm, ph, bc, pbc_h, [port_h], 0)
    A_o, b_o = orc.assemble_maxwell_periodic(om, po, pec, pbc_o, [port_o], 0)
    rp, ci, va = asm.A.to_csr()
    A_h = sp.csr_matrix((va, ci, rp), shape=asm.A.shape)
    assert abs(A_h - A_o).max() <= 1e-12 * abs(A_o).max()
    # pattern after sparseView(): identical up to entries that are rounding noise in one summation order
    d = (abs(A_h) > 1e-9 * abs(A_o).max()).astype(int) - (abs(A_o) > 1e-9 * abs(A_o).max()).astype(int)
    assert abs(d).sum() == 0
    assert np.max(np.abs(asm.b.to_numpy() - b_o)) <= 1e-13 * np.max(np.abs(b_o))
    # a non-real Bloch phase makes A neither symmetric nor Hermitian => BiCGSTAB on the device
    S_h = pe.calculate_sparams_periodic(hm, ph, bc, pbc_h, [port_h])
    S_o = orc.calculate_sparams_periodic(om, po, pec, pbc_o, [port_o])
    assert np.max(np.abs(S_h - S_o)) <= 1e-6 * max(1.0, np.max(np.abs(S_o)))
    assert abs(S_h[0, 0]) <= 1.0 + 1e-6
