"""SURVEY 8f rows f1 (nodal port eigenmodes, port-face extraction, modal line-integral ports) and f4 (Touchstone
writers): C++ host layer (pyedgefem) against the oracle restatement and against analytic known answers.  CPU only."""
import math
import os

import numpy as np
import pytest

import edgefem_oracle as orc
from conftest import GOLDEN, load_fixture_mesh
from edgefem_b200 import load_pyedgefem

pe = load_pyedgefem()
A_WR90, B_WR90 = 0.02286, 0.01016


def host_mesh(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return pe.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], z["node_ids"].tolist())


@pytest.fixture(scope="module")
def wr90():
    return host_mesh("rect_waveguide"), load_fixture_mesh("rect_waveguide")


def test_extract_surface_mesh(wr90):
    hm, om = wr90
    for tag in (2, 3):
        hs, os_ = pe.extract_surface_mesh(hm, tag), orc.extract_surface_mesh(om, tag)
        assert list(hs.volume_tri_indices) == os_.volume_tri_indices.tolist()
        assert [n.id for n in hs.mesh.nodes] == os_.node_ids.tolist()  # first-seen order over the face's tris
        xy = np.array([[n.xyz[0], n.xyz[1], n.xyz[2]] for n in hs.mesh.nodes])
        assert np.array_equal(xy[:, :2], os_.xy) and np.all(xy[:, 2] == 0.0)  # z dropped
        assert [list(t.conn)[:3] for t in hs.mesh.tris] == os_.tri_conn.tolist()
        bl = hs.mesh.boundary_lines_array()
        assert sorted(map(tuple, bl[:, :2].tolist())) == [tuple(r) for r in os_.boundary_lines.tolist()]
        assert np.all(bl[:, 2] == 1)  # boundary of the face = PEC by default
    assert len(pe.extract_surface_mesh(hm, 999).mesh.tris) == 0


@pytest.mark.parametrize("pol,tm", [("TE", False), ("TM", True)])
def test_solve_port_eigens_vs_oracle_and_analytic(wr90, pol, tm):
    hm, om = wr90
    hs, os_ = pe.extract_surface_mesh(hm, 2), orc.extract_surface_mesh(om, 2)
    omega = 2 * math.pi * 10e9
    hmodes = pe.solve_port_eigens(hs.mesh, 3, omega, 1.0, 1.0, getattr(pe.ModePolarization, pol))
    omodes = orc.solve_port_eigens(os_, 3, omega, 1.0, 1.0, tm=tm)
    assert len(hmodes) == len(omodes) == 3
    for h, (o, fld) in zip(hmodes, omodes):
        assert abs(h.kc - o.kc) / o.kc < 1e-9 and abs(h.fc - o.fc) / o.fc < 1e-9
        assert abs(h.beta - o.beta) <= 1e-9 * max(1.0, abs(o.beta))
        assert abs(h.Z0 - o.Z0) <= 1e-9 * max(1.0, abs(o.Z0))
        assert h.omega == omega and abs(h.mu - orc.MU0) < 1e-20 and abs(h.eps - orc.EPS0) < 1e-24
        hf = np.asarray(h.field)
        assert np.max(np.abs(hf - fld)) <= 1e-7 * np.max(np.abs(fld))  # same sign convention, unit modal power
    # analytic cut-offs of the a x b rectangle (P1 elements on this face: a few per cent)
    if not tm:
        assert abs(hmodes[0].kc - math.pi / A_WR90) / (math.pi / A_WR90) < 0.02  # TE10
        assert abs(hmodes[0].fc - 6.557e9) / 6.557e9 < 0.02
    else:
        k11 = math.hypot(math.pi / A_WR90, math.pi / B_WR90)
        assert abs(hmodes[0].kc - k11) / k11 < 0.06  # TM11
        assert hmodes[0].beta.imag != 0.0 or hmodes[0].beta.real == 0.0  # TM11 is cut off at 10 GHz: beta on the imaginary axis
    with pytest.raises(RuntimeError):
        pe.solve_port_eigens(hs.mesh, 1, 0.0, 1.0, 1.0, pe.ModePolarization.TE)
    with pytest.raises(RuntimeError):
        pe.solve_port_eigens(hm.__class__(), 1, omega, 1.0, 1.0, pe.ModePolarization.TE)


def test_build_wave_port_te10_and_nodal_mode(wr90):
    hm, om = wr90
    hs, os_ = pe.extract_surface_mesh(hm, 2), orc.extract_surface_mesh(om, 2)
    # analytic TE10 field sampled on the face
    hmode = pe.solve_te10_mode(pe.RectWaveguidePort(A_WR90, B_WR90), 10e9)
    omode = orc.solve_te10_mode(A_WR90, B_WR90, 10e9)
    pe.populate_te10_field(hs, pe.RectWaveguidePort(A_WR90, B_WR90), hmode)
    ofld = orc.populate_te10_field(os_, A_WR90, B_WR90, omode)
    assert np.max(np.abs(np.asarray(hmode.field) - ofld)) <= 1e-13 * np.max(np.abs(ofld))
    hp = pe.build_wave_port(hm, hs, hmode)
    op = orc.build_wave_port(om, os_, omode, ofld)
    assert hp.surface_tag == 2 and list(hp.edges) == op.edges
    assert np.max(np.abs(np.asarray(hp.weights) - op.weights)) <= 1e-12 * np.max(np.abs(op.weights))
    # the nodal FEM mode gives (up to the discretisation error) the same port
    hm1 = pe.solve_port_eigens(hs.mesh, 1, 2 * math.pi * 10e9, 1.0, 1.0, pe.ModePolarization.TE)[0]
    hp1 = pe.build_wave_port(hm, hs, hm1)
    w0, w1 = np.asarray(hp.weights), np.asarray(hp1.weights)
    cos = abs(np.vdot(w0, w1)) / (np.linalg.norm(w0) * np.linalg.norm(w1))
    assert cos > 0.995
    bad = pe.PortMode()
    bad.field = np.zeros(3, dtype=complex)
    with pytest.raises(RuntimeError):
        pe.build_wave_port(hm, hs, bad)


def test_build_wave_port_from_eigenvector(wr90):
    hm, om = wr90
    hs, os_ = pe.extract_surface_mesh(hm, 3), orc.extract_surface_mesh(om, 3)
    pec = orc.build_edge_pec(om, 1)
    rng = np.random.default_rng(3)
    ev = rng.standard_normal(om.num_edges)
    hmode = pe.solve_te10_mode(pe.RectWaveguidePort(A_WR90, B_WR90), 10e9)
    omode = orc.solve_te10_mode(A_WR90, B_WR90, 10e9)
    hp = pe.build_wave_port_from_eigenvector(hm, hs, ev, hmode, set(pec))
    op = orc.build_wave_port_from_eigenvector(om, os_, ev, omode, pec)
    assert list(hp.edges) == op.edges
    w = np.asarray(hp.weights)
    assert np.max(np.abs(w - op.weights)) <= 1e-13 * np.max(np.abs(op.weights))
    assert np.all(w.real == 0.0)  # j * real eigenvector
    assert abs(np.vdot(w, w).real - math.sqrt(hmode.Z0.real)) < 1e-9  # ||w||^2 = sqrt(Z0)


def test_straight_waveguide_sparams():
    port = pe.RectWaveguidePort(A_WR90, B_WR90)
    s = pe.straight_waveguide_sparams(port, 0.05, 10e9)
    o = orc.straight_waveguide_sparams(A_WR90, 0.05, 10e9)
    assert (s.s11, s.s21, s.s12, s.s22) == pytest.approx(o, abs=1e-14)
    assert abs(abs(s.s21) - 1.0) < 1e-14 and s.s11 == 0
    below = pe.straight_waveguide_sparams(port, 0.05, 5e9)  # below the 6.56 GHz cut-off: full reflection
    assert (below.s11, below.s21, below.s12, below.s22) == (1, 0, 0, 1)


@pytest.mark.parametrize("fmt", ["RI", "MA", "DB"])
@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_write_touchstone_nport_text(tmp_path, fmt, n):
    rng = np.random.default_rng(10 * n + len(fmt))
    freq = [8e9, 9.123456789012e9, 1.2e10]
    S = [rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) for _ in freq]
    S[0][0, 0] = 0.0  # exercises the 1e-20 floor of the dB format
    opts = pe.TouchstoneOptions()
    opts.format = getattr(pe.TouchstoneFormat, fmt)
    opts.z0 = 75.0
    path = str(tmp_path / ("x" + pe.touchstone_extension(n)))
    pe.write_touchstone_nport(path, freq, S, opts)
    assert open(path).read() == orc.touchstone_nport_text(freq, S, fmt, 75.0)
    assert path.endswith(".s%dp" % n)


def test_write_touchstone_legacy_and_errors(tmp_path):
    freq = [1e9, 2.5e9]
    sp = []
    for f in freq:
        s = pe.straight_waveguide_sparams(pe.RectWaveguidePort(0.3, 0.1), 0.2, f)
        sp.append(s)
    path = str(tmp_path / "legacy.s2p")
    pe.write_touchstone(path, freq, sp)
    want = orc.touchstone_legacy_text(freq, [(s.s11, s.s21, s.s12, s.s22) for s in sp])
    assert open(path).read() == want and want.startswith("# Hz S RI R 50\n")
    with pytest.raises(RuntimeError):
        pe.write_touchstone_nport(str(tmp_path / "e.s2p"), [], [])
    with pytest.raises(RuntimeError):
        pe.write_touchstone_nport(str(tmp_path / "e.s2p"), [1e9, 2e9], [np.eye(2, dtype=complex)])
    with pytest.raises(RuntimeError):
        pe.write_touchstone_nport(str(tmp_path / "e.s2p"), [1e9, 2e9], [np.eye(2, dtype=complex), np.eye(3, dtype=complex)])
    with pytest.raises(RuntimeError):
        pe.touchstone_extension(0)
    with pytest.raises(RuntimeError):
        pe.touchstone_extension(100)
    assert pe.touchstone_extension(12) == ".s12p"
