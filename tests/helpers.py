"""Test-side glue: turn oracle objects into C-ABI inputs (tests may import the oracle)."""
import math

import numpy as np

import edgefem_oracle as orc
from edgefem_b200 import cabi


def device_mesh(ctx, mesh: orc.Mesh) -> cabi.DeviceMesh:
    tet_nodes = mesh.node_idx_of(mesh.tet_conn).astype(np.int32)
    edge_nodes = mesh.node_idx_of(mesh.edges).astype(np.int32)
    return cabi.DeviceMesh(ctx, mesh.xyz, tet_nodes, mesh.tet_edges, mesh.tet_orient.astype(np.int8), mesh.tet_phys, edge_nodes)


def pec_flags(mesh: orc.Mesh, pec) -> np.ndarray:
    return orc.pec_mask(mesh, pec).astype(np.uint8)


def csr_of(sysd: cabi.DeviceSystem, matrix=0):
    import scipy.sparse as sp

    rp, ci = sysd.pattern()
    return sp.csr_matrix((sysd.values(matrix), ci, rp), shape=(sysd.m, sysd.m))


def rel_entry_err(v_gpu: np.ndarray, v_ref: np.ndarray) -> float:
    """max |a-b| / max(|b|, floor) with floor = 1e-3 * median |b| over non-zero entries: entries that are
    structurally tiny (sums that cancel) are compared against the scale of their neighbours."""
    nz = np.abs(v_ref)[np.abs(v_ref) > 0]
    floor = 1e-3 * (np.median(nz) if nz.size else 1.0)
    return float(np.max(np.abs(v_gpu - v_ref) / np.maximum(np.abs(v_ref), floor)))


def row_rel_err(A_gpu, A_ref) -> float:
    """max_ij |gpu_ij - ref_ij| / max_j |ref_ij|  (error relative to the largest entry of the row).
    This is the asserted 1e-12 fp64 bar: an entry that is itself the result of cancellation -- inside
    c_i.c_j for near-orthogonal curls, or across the tets sharing the edge pair -- cannot match to
    1e-12 of its OWN magnitude between two correct implementations with different operation order
    (numpy vs FMA-contracted CUDA vs Eigen's setFromTriplets order); on the scale of its row it does.
    Both matrices share one CSR pattern (asserted by the callers)."""
    import scipy.sparse as sp

    A_gpu, A_ref = sp.csr_matrix(A_gpu), sp.csr_matrix(A_ref)
    m = A_ref.shape[0]
    rows = np.repeat(np.arange(m), np.diff(A_ref.indptr))
    rmax = np.zeros(m)
    np.maximum.at(rmax, rows, np.abs(A_ref.data))
    rmax = np.maximum(rmax, np.finfo(float).tiny)
    return float(np.max(np.abs(A_gpu.data - A_ref.data) / rmax[rows]))


def sum_rel_err(A_gpu, A_ref, scale) -> float:
    """max |gpu - ref| / sum_of_abs_contributions, entrywise; A_gpu and A_ref share one CSR pattern
    (asserted by the callers).  `scale` (orc.volume_abs_scale, any pattern) is looked up per entry.
    This is the 1e-12 fp64 bar of north_star: an entry whose contributions cancel cannot be
    reproduced to 1e-12 of its OWN magnitude by any summation order, Eigen's included."""
    import scipy.sparse as sp

    A_gpu, A_ref, sc = sp.csr_matrix(A_gpu), sp.csr_matrix(A_ref), sp.csr_matrix(scale)
    sc.sort_indices()
    m = A_ref.shape[0]
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(A_ref.indptr))
    keys = rows * m + A_ref.indices
    srows = np.repeat(np.arange(m, dtype=np.int64), np.diff(sc.indptr))
    skeys = srows * m + sc.indices
    idx = np.minimum(np.searchsorted(skeys, keys), max(0, skeys.size - 1))
    s = np.where(skeys[idx] == keys, np.abs(sc.data[idx]), 0.0) if skeys.size else np.zeros(keys.size)
    # entries whose contributions are all exactly 0 in one summation (right angles in structured
    # regions) but rounding noise in the other (FMA contraction) are measured on the matrix scale
    s = np.maximum(s, 1e-6 * np.abs(sc.data).max())
    return float(np.max(np.abs(A_gpu.data - A_ref.data) / s))


def port_device(sysd, mesh, pec, port: orc.WavePort, with_mass=True) -> cabi.DevicePort:
    if with_mass:
        ms = orc.assemble_port_surface_mass(mesh, port.surface_tag, pec).tocoo()
        return cabi.DevicePort(sysd, port.edges, port.weights, ms.row, ms.col, ms.data)
    return cabi.DevicePort(sysd, port.edges, port.weights)


def eigenmode_sweep_gpu(ctx, mesh, pec, ports, freqs, port_abc_scale=1.0, tol=1e-10, precond=cabi.PRECOND_AUX, method=cabi.METHOD_AUTO,
                        keep=None):
    """calculate_sparams_eigenmode (src/assemble_maxwell.cpp:637-787) for a batch of frequencies through the C-ABI."""
    dm = device_mesh(ctx, mesh)
    pe = np.nonzero(orc.pec_mask(mesh, pec))[0].astype(np.int32)
    F, P = len(freqs), len(ports)
    sysd = cabi.DeviceSystem.from_mesh(dm, pe, pe, n_matrix=F, n_rhs=P)
    sysd.set_dirichlet(pec_flags(mesh, pec))
    mats, keepalive = cabi.make_materials(len(dm.slot_tags))
    omegas = [2 * math.pi * f for f in freqs]
    sysd.assemble_volume(omegas, mats)
    dports = [port_device(sysd, mesh, pec, p) for p in ports]
    for dp in dports:
        dp.normalize_mass()
    betas = np.zeros((F, P), dtype=np.complex128)
    for fi, om in enumerate(omegas):
        b = orc.port_betas(mesh, orc.MaxwellParams(omega=om), ports)
        betas[fi] = b
    for pi, dp in enumerate(dports):
        dp.add_mass(1j * port_abc_scale * betas[:, pi])
    for fi in range(F):
        for a in range(P):
            dports[a].rhs_mass(fi * P + a, 2.0 * 1j * port_abc_scale * betas[fi, a])
    res = sysd.solve(method=method, precond=precond, tol=tol, symmetric=True)
    S = np.zeros((F, P, P), dtype=np.complex128)
    for fi in range(F):
        for a in range(P):
            for j in range(P):
                v = dports[j].project_mass(fi * P + a)
                S[fi, j, a] = v - 1.0 if j == a else v
    if keep is not None:
        keep.update(sys=sysd, ports=dports, mesh=dm, mats=(mats, keepalive))
    else:
        for dp in dports:
            dp.close()
        sysd.close()
        dm.close()
    return S, res
