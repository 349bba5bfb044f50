"""SURVEY 8f-f3 on the device: Huygens-surface extraction (efb_huygens_eval) and the near-to-far-field transformation
(efb_stratton_chu) against the oracle restatements and against the analytic Hertzian-dipole far field
(the reference's own check, tests/test_ntf.cpp)."""
import math

import numpy as np
import pytest

import edgefem_oracle as orc
from conftest import load_fixture_mesh
from edgefem_b200 import cabi, load_pyedgefem

pytestmark = pytest.mark.gpu
pe = load_pyedgefem()


def host_mesh(om):
    return pe.mesh_from_arrays(om.xyz, om.tet_conn, om.tet_phys, om.tri_conn, om.tri_phys, om.node_ids.tolist())


def test_evaluate_edge_field_helpers_cpu_side():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((4, 3))
    p = X.mean(axis=0) + 0.05 * rng.standard_normal(3)
    dofs = rng.standard_normal(6) + 1j * rng.standard_normal(6)
    orient = [1, -1, 1, 1, -1, -1]
    verts = [tuple(v) for v in X]
    assert np.allclose(pe.compute_barycentric(verts, tuple(p)), orc.compute_barycentric(X, p), rtol=0, atol=1e-13)
    assert np.allclose(pe.whitney_edge_curls(verts), orc.whitney_edge_curls(X), rtol=1e-13, atol=1e-13)
    assert np.allclose(pe.evaluate_edge_field(verts, orient, dofs, tuple(p)), orc.evaluate_edge_field(X, orient, dofs, p), rtol=1e-12, atol=1e-13)
    # tangential continuity property of Whitney elements: the line integral of W_e along edge e is 1, along the others 0
    for e, (a, b) in enumerate(orc._EDGE_PAIRS):
        unit = np.zeros(6, dtype=complex)
        unit[e] = 1.0
        for e2, (a2, b2) in enumerate(orc._EDGE_PAIRS):
            mid = 0.5 * (X[a2] + X[b2])
            val = np.asarray(pe.evaluate_edge_field(verts, [1] * 6, unit, tuple(mid))) @ (X[b2] - X[a2])
            assert abs(val - (1.0 if e == e2 else 0.0)) < 1e-12


@pytest.mark.parametrize("tag", [2, 3, 1])
def test_extract_huygens_surface_vs_oracle(tag):
    """Random solution vector on the WR-90 mesh: every tagged triangle's r, n, area, E_tan, H_tan (ports and the PEC wall)."""
    om = load_fixture_mesh("rect_waveguide")
    hm = host_mesh(om)
    rng = np.random.default_rng(tag)
    x = rng.standard_normal(om.num_edges) + 1j * rng.standard_normal(om.num_edges)
    omega, mu_r = 2 * math.pi * 10e9, 1.3 - 0.2j
    d = pe.extract_huygens_surface(hm, pe.VecC(x), tag, omega, mu_r)
    o = orc.extract_huygens_surface(om, x, tag, omega, mu_r)
    assert d.r.shape == o["r"].shape and d.r.shape[0] == int((om.tri_phys == tag).sum())
    assert np.allclose(d.r, o["r"], rtol=0, atol=1e-15) and np.allclose(d.n, o["n"], rtol=0, atol=1e-13)
    assert np.allclose(d.area, o["area"], rtol=1e-13, atol=0)
    for a, b in ((d.E_tan, o["E_tan"]), (d.H_tan, o["H_tan"])):
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
    with pytest.raises(RuntimeError):
        pe.extract_huygens_surface(hm, pe.VecC(x), 999, omega)


def dipole_surface(k0, R=0.5, nth=24, nph=48):
    """z-directed Hertzian dipole (I l = 1) at the origin sampled on a sphere: exact near fields in Balanis's
    e^{+jwt} form, conjugated wholesale to the kernel's e^{-jwt} convention like tests/test_ntf.cpp:96-113."""
    th = (np.arange(nth) + 0.5) * np.pi / nth
    ph = (np.arange(nph) + 0.5) * 2 * np.pi / nph
    T, P = np.meshgrid(th, ph, indexing="ij")
    T, P = T.ravel(), P.ravel()
    rhat = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=1)
    that = np.stack([np.cos(T) * np.cos(P), np.cos(T) * np.sin(P), -np.sin(T)], axis=1)
    phat = np.stack([-np.sin(P), np.cos(P), np.zeros_like(P)], axis=1)
    kr = k0 * R
    g = np.exp(-1j * kr) / (4 * np.pi * R)
    eta = orc.Z0_FREE
    Er = eta * 2 * np.cos(T) * g * (1 / R) * (1 + 1 / (1j * kr))
    Et = 1j * eta * k0 * np.sin(T) * g * (1 + 1 / (1j * kr) - 1 / kr ** 2)
    Hp = 1j * k0 * np.sin(T) * g * (1 + 1 / (1j * kr))
    E = Er[:, None] * rhat + Et[:, None] * that
    H = Hp[:, None] * phat
    area = R * R * np.sin(T) * (np.pi / nth) * (2 * np.pi / nph)
    return R * rhat, rhat, np.conj(E), np.conj(H), area


def test_stratton_chu_vs_oracle_and_hertzian_dipole():
    k0 = 2 * np.pi / 0.3
    r, n, E, H, area = dipole_surface(k0)
    theta = np.linspace(0.0, np.pi, 19)
    phi = np.linspace(0.0, 2 * np.pi, 13)
    ctx = cabi.Ctx(0)
    TH, PH = np.meshgrid(theta, phi, indexing="ij")
    et, ep = cabi.stratton_chu(ctx, r, n, E, H, area, TH.ravel(), PH.ravel(), k0)
    ot, op = orc.stratton_chu(r, n, E, H, area, TH.ravel(), PH.ravel(), k0)
    scale = np.max(np.abs(ot))
    assert np.max(np.abs(et - ot)) <= 1e-12 * scale and np.max(np.abs(ep - op)) <= 1e-12 * scale
    ctx.close()
    # analytic r-normalised far field of the dipole: |E_theta| = eta k0 sin(theta) / (4 pi), E_phi = 0
    # (the reference's own acceptance test, tests/test_ntf.cpp:129-139: 2 % quadrature tolerance)
    want = orc.Z0_FREE * k0 * np.sin(TH.ravel()) / (4 * np.pi)
    assert np.max(np.abs(np.abs(et) - want)) <= 0.02 * np.max(want)
    assert np.max(np.abs(ep)) <= 1e-3 * np.max(want)
    # pyedgefem API on top of the same kernel
    pat = pe.stratton_chu_3d(r, n, E, H, list(area), list(theta), list(phi), k0)
    assert np.allclose(pat.E_theta, et.reshape(TH.shape), rtol=0, atol=1e-13 * scale)
    assert np.allclose(pat.theta_grid, TH) and np.allclose(pat.phi_grid, PH)
    D = pe.compute_directivity(pat)
    assert abs(D - orc.compute_directivity(theta, phi, np.asarray(pat.E_theta), np.asarray(pat.E_phi))) < 1e-10
    assert abs(D - 1.5) < 0.05  # Hertzian dipole
    assert pe.compute_max_gain(pat, 0.5) == pytest.approx(0.5 * D)
    e_bw, h_bw = pe.compute_hpbw(pat)
    assert 80.0 <= e_bw <= 110.0  # 90 degrees for sin^2(theta), on a 10-degree grid
    cut = pe.stratton_chu_2d(r, n, E, H, list(area), list(theta), 0.3, k0)
    o2t, o2p = orc.stratton_chu(r, n, E, H, area, theta, np.full_like(theta, 0.3), k0)
    assert np.allclose([c.e_theta for c in cut], o2t, rtol=0, atol=1e-12 * scale)
    assert np.allclose([c.theta_deg for c in cut], np.degrees(theta))
    assert np.all(pat.pattern_dB() <= 1e-12) and pat.power_pattern().shape == TH.shape
