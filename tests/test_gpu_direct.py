"""GPU: dense LU with partial pivoting (efb_solve_direct), the robust fallback behind solve_linear's contract
(src/solver.cpp:11-33 `use_direct`, :55-80 "->SparseLU"), against SuperLU."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import edgefem_oracle as orc
import helpers as H
from edgefem_b200 import cabi, load_pyedgefem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Ctx(0)
    yield c
    c.close()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 200, 1000])
def test_direct_random_nonsymmetric(ctx, n):
    """Generic CSR systems (no Dirichlet flags), non-symmetric, rows that NEED pivoting (zero diagonal), two rhs; ragged
    sizes around the block width 32."""
    rng = np.random.default_rng(n)
    A = sp.random(n, n, density=min(1.0, 8.0 / n), random_state=rng.integers(1 << 30), format="lil", dtype=np.float64).astype(complex)
    A = A + 1j * sp.random(n, n, density=min(1.0, 8.0 / n), random_state=rng.integers(1 << 30), format="lil")
    perm = rng.permutation(n)
    A = sp.lil_matrix(A)
    for i in range(n):
        A[i, perm[i]] = 3.0 + rng.standard_normal() + 1j * rng.standard_normal()  # a strong entry OFF the diagonal in general
        if perm[i] != i:
            A[i, i] = 0.0
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    A.sort_indices()
    sysd = cabi.DeviceSystem.from_csr(ctx, A.indptr, A.indices, A.data, n_matrix=1, n_rhs=2)
    bs = [rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(2)]
    for k in range(2):
        sysd.rhs_set(k, bs[k])
    res = sysd.solve_direct()
    lu = spla.splu(sp.csc_matrix(A))
    for k in range(2):
        x = sysd.x_get(k)
        xr = lu.solve(bs[k])
        assert res[k]["converged"] and res[k]["method"] == cabi.METHOD_DIRECT, res[k]
        assert np.linalg.norm(x - xr) <= 1e-9 * np.linalg.norm(xr), (n, k, res[k])
    sysd.close()


def test_direct_wr90_eigenmode_system(ctx, wr90):
    """The WR-90 eigenmode system (Dirichlet rows dropped from the factorisation, both ports as right-hand sides)."""
    mesh, pec = wr90
    ports = orc.wr90_ports(mesh, pec, 10e9)
    dm = H.device_mesh(ctx, mesh)
    pe_idx = np.nonzero(orc.pec_mask(mesh, pec))[0].astype(np.int32)
    sysd = cabi.DeviceSystem.from_mesh(dm, pe_idx, pe_idx, n_matrix=1, n_rhs=2)
    sysd.set_dirichlet(H.pec_flags(mesh, pec))
    mats, keepalive = cabi.make_materials(len(dm.slot_tags))
    om = 2 * math.pi * 10e9
    sysd.assemble_volume([om], mats)
    dports = [H.port_device(sysd, mesh, pec, p) for p in ports]
    for dp in dports:
        dp.normalize_mass()
    betas = orc.port_betas(mesh, orc.MaxwellParams(omega=om), ports)
    for i, dp in enumerate(dports):
        dp.add_mass(np.array([1j * betas[i]]))
    for a in range(2):
        dports[a].rhs_mass(a, 2.0 * 1j * betas[a])
    res = sysd.solve_direct()
    A = H.csr_of(sysd)
    lu = spla.splu(sp.csc_matrix(A))
    for a in range(2):
        b = sysd.rhs_get(a)
        x = sysd.x_get(a)
        assert res[a]["converged"] and res[a]["residual"] < 1e-10, res[a]
        xr = lu.solve(b)
        assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)
    S = np.zeros((2, 2), dtype=complex)
    for j, dp in enumerate(dports):
        for a in range(2):
            S[j, a] = dp.project_mass(a) - (1.0 if j == a else 0.0)
    S_o = orc.wr90_sparams(mesh, pec, 10e9, ports)
    assert np.max(np.abs(S - S_o)) <= 1e-6
    sysd.close()
    dm.close()
