"""CPU oracle: a numpy/scipy restatement of EdgeFEM's frequency-domain solve hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``edgefem_b200``: C++ host + CUDA behind the
C-ABI in ``include/edgefem_b200.h``) never does.

Why a restatement: the reference (jman4162/EdgeFEM, C++20) hard-requires Eigen >= 3.4,
which is neither vendored under /root/reference nor installed in this image (no
network), so the reference cannot be compiled here (``oracle/_ref`` does not exist).
The arithmetic of the assembly path is fully determined by the reference sources; the
only third-party arithmetic is Eigen 3.4.0's ``setFromTriplets`` duplicate summation
(order unspecified => value parity is 1e-12 relative, not bitwise), 3x3 inverse /
determinant, and the sparse solvers.  Solves here use SuperLU (``scipy.sparse.linalg.splu``)
-- the code base Eigen's ``SparseLU`` derives from -- and are residual-checked.

Parity pinning: ``tests/test_oracle_kat.py`` checks this oracle against the reference's
published WR-90 table (docs/validation.md:22-27), the alpha sweep (docs/validation.md:51-57),
the enforced thresholds of tests/benchmark_wr90.cpp:91, tests/test_eigenmode_sparams.cpp:116,
tests/test_cavity_eigenmodes.cpp:276 and the structural invariants of
tests/test_edge_indexing.cpp, tests/test_maxwell.cpp, tests/test_triangle_mass_matrix.cpp,
tests/test_dispersive_materials.cpp, tests/test_periodic.cpp.  Paths with no enforced
upstream known answer (periodic elimination, dispersive materials inside a solve,
lumped-port S11, frequency_sweep outputs, PML values) are "parity unpinned" upstream:
for those the GPU path is compared with this restatement only.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# src/assemble_maxwell.cpp:39-43
C0 = 299792458.0
MU0 = 4.0 * math.pi * 1e-7
EPS0 = 1.0 / (MU0 * C0 * C0)
ETA0 = MU0 * C0

TET_PAIRS = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])  # src/mesh_gmsh.cpp:105-108
TRI_PAIRS = np.array([[0, 1], [1, 2], [2, 0]])  # src/mesh_gmsh.cpp:109-111


# --------------------------------------------------------------------------------------
# Mesh + edge numbering  (include/edgefem/mesh.hpp:13-60, src/mesh_gmsh.cpp:15-146)
# --------------------------------------------------------------------------------------
@dataclass
class Mesh:
    node_ids: np.ndarray  # int64 [n]
    xyz: np.ndarray  # float64 [n,3]
    tet_conn: np.ndarray  # int64 [t,4]  node *ids*
    tet_phys: np.ndarray  # int32 [t]
    tri_conn: np.ndarray  # int64 [k,3]
    tri_phys: np.ndarray  # int32 [k]
    tet_edges: np.ndarray = None  # int32 [t,6]
    tet_orient: np.ndarray = None  # int32 [t,6]
    tri_edges: np.ndarray = None  # int32 [k,3]
    tri_orient: np.ndarray = None  # int32 [k,3]
    edges: np.ndarray = None  # int64 [m,2]  (n0 < n1) node ids
    node_index: Dict[int, int] = None

    @property
    def num_edges(self) -> int:
        return int(self.edges.shape[0])

    def tet_xyz(self) -> np.ndarray:
        """[t,4,3] vertex coordinates (nodeIndex lookup, src/assemble_maxwell.cpp:115-119)."""
        idx = self.node_idx_of(self.tet_conn)
        return self.xyz[idx]

    def tri_xyz(self) -> np.ndarray:
        idx = self.node_idx_of(self.tri_conn)
        return self.xyz[idx]

    def node_idx_of(self, ids: np.ndarray) -> np.ndarray:
        ids = np.asarray(ids)
        if ids.size == 0:
            return ids.astype(np.int64)
        lut = self._lut()
        return lut[ids]

    def _lut(self) -> np.ndarray:
        if getattr(self, "_lut_cache", None) is None:
            mx = int(self.node_ids.max()) if self.node_ids.size else 0
            lut = np.full(mx + 1, -1, dtype=np.int64)
            # later duplicates overwrite earlier ones like mesh.nodeIndex[id] = ... (mesh_gmsh.cpp:31)
            lut[self.node_ids] = np.arange(self.node_ids.size)
            self._lut_cache = lut
        return self._lut_cache


def make_edge_key(a: int, b: int) -> int:
    """include/edgefem/mesh.hpp:56-60."""
    if a > b:
        a, b = b, a
    return ((a << 32) ^ b) & 0xFFFFFFFFFFFFFFFF


def build_edges(mesh: Mesh) -> None:
    """Global edge numbering, first-seen order over tets then tris (src/mesh_gmsh.cpp:104-146).

    Pure-python dict walk: the literal restatement (used for small meshes and as the
    reference for the vectorised variant below)."""
    edge_index: Dict[int, int] = {}
    edges: List[Tuple[int, int]] = []

    def walk(conn: np.ndarray, pairs: np.ndarray):
        ne = pairs.shape[0]
        out_e = np.zeros((conn.shape[0], ne), dtype=np.int32)
        out_o = np.zeros((conn.shape[0], ne), dtype=np.int32)
        for t in range(conn.shape[0]):
            c = conn[t]
            for e in range(ne):
                a = int(c[pairs[e, 0]])
                b = int(c[pairs[e, 1]])
                sign = 1 if a < b else -1
                key = make_edge_key(a, b)
                idx = edge_index.get(key)
                if idx is None:
                    idx = len(edges)
                    edge_index[key] = idx
                    edges.append((min(a, b), max(a, b)))
                out_e[t, e] = idx
                out_o[t, e] = sign
        return out_e, out_o

    mesh.tet_edges, mesh.tet_orient = walk(mesh.tet_conn, TET_PAIRS)
    mesh.tri_edges, mesh.tri_orient = walk(mesh.tri_conn, TRI_PAIRS)
    mesh.edges = np.array(edges, dtype=np.int64).reshape(-1, 2)


def build_edges_fast(mesh: Mesh) -> None:
    """Vectorised equivalent of build_edges (same numbering; used for big synthetic meshes).

    First-seen order == order of the first occurrence of each key in the flattened
    (tets row-major, then tris row-major) slot sequence."""
    ta = mesh.tet_conn[:, TET_PAIRS[:, 0]].reshape(-1)
    tb = mesh.tet_conn[:, TET_PAIRS[:, 1]].reshape(-1)
    ra = mesh.tri_conn[:, TRI_PAIRS[:, 0]].reshape(-1) if mesh.tri_conn.size else np.zeros(0, np.int64)
    rb = mesh.tri_conn[:, TRI_PAIRS[:, 1]].reshape(-1) if mesh.tri_conn.size else np.zeros(0, np.int64)
    a = np.concatenate([ta, ra]).astype(np.int64)
    b = np.concatenate([tb, rb]).astype(np.int64)
    lo = np.minimum(a, b)
    hi = np.maximum(a, b)
    key = (lo << 32) ^ hi  # ids < 2^31 in every mesh we build => injective
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique keys sorted by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    eid = rank[inv].astype(np.int32)
    sign = np.where(a < b, 1, -1).astype(np.int32)
    nt = mesh.tet_conn.shape[0] * 6
    mesh.tet_edges = eid[:nt].reshape(-1, 6)
    mesh.tet_orient = sign[:nt].reshape(-1, 6)
    mesh.tri_edges = eid[nt:].reshape(-1, 3)
    mesh.tri_orient = sign[nt:].reshape(-1, 3)
    fo = first[order]
    mesh.edges = np.stack([lo[fo], hi[fo]], axis=1).astype(np.int64)


def load_gmsh_v2(path: str, fast: bool = False) -> Mesh:
    """Gmsh v2 ASCII reader (src/mesh_gmsh.cpp:15-75): first tag = phys, types 2/1/4 kept."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    i = 0
    node_ids: List[int] = []
    xyz: List[Tuple[float, float, float]] = []
    tets: List[Tuple[int, int, int, int]] = []
    tet_phys: List[int] = []
    tris: List[Tuple[int, int, int]] = []
    tri_phys: List[int] = []
    opened = False
    while i < len(lines):
        line = lines[i].rstrip("\r")
        opened = True
        if line == "$Nodes":
            n = int(lines[i + 1].split()[0])
            for k in range(n):
                p = lines[i + 2 + k].split()
                node_ids.append(int(p[0]))
                xyz.append((float(p[1]), float(p[2]), float(p[3])))
            i += 2 + n
        elif line == "$Elements":
            m = int(lines[i + 1].split()[0])
            for k in range(m):
                p = lines[i + 2 + k].split()
                etype = int(p[1])
                ntags = int(p[2])
                phys = int(p[3]) if ntags > 0 else 0
                c = p[3 + ntags:]
                if etype == 2:
                    tris.append((int(c[0]), int(c[1]), int(c[2])))
                    tri_phys.append(phys)
                elif etype == 4:
                    tets.append((int(c[0]), int(c[1]), int(c[2]), int(c[3])))
                    tet_phys.append(phys)
            i += 2 + m
        else:
            i += 1
    if not opened:
        raise RuntimeError("Failed to open mesh file: " + path)
    mesh = Mesh(
        node_ids=np.array(node_ids, dtype=np.int64),
        xyz=np.array(xyz, dtype=np.float64).reshape(-1, 3),
        tet_conn=np.array(tets, dtype=np.int64).reshape(-1, 4),
        tet_phys=np.array(tet_phys, dtype=np.int32),
        tri_conn=np.array(tris, dtype=np.int64).reshape(-1, 3),
        tri_phys=np.array(tri_phys, dtype=np.int32),
    )
    (build_edges_fast if fast else build_edges)(mesh)
    return mesh


def mesh_from_arrays(xyz, tets, tet_phys, tris, tri_phys, node_ids=None, fast=True) -> Mesh:
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    if node_ids is None:
        node_ids = np.arange(1, xyz.shape[0] + 1, dtype=np.int64)
    mesh = Mesh(
        node_ids=np.asarray(node_ids, dtype=np.int64),
        xyz=xyz,
        tet_conn=np.asarray(tets, dtype=np.int64).reshape(-1, 4),
        tet_phys=np.asarray(tet_phys, dtype=np.int32),
        tri_conn=np.asarray(tris, dtype=np.int64).reshape(-1, 3),
        tri_phys=np.asarray(tri_phys, dtype=np.int32),
    )
    (build_edges_fast if fast else build_edges)(mesh)
    return mesh


# --------------------------------------------------------------------------------------
# PEC  (src/bc.cpp:47-80)
# --------------------------------------------------------------------------------------
def build_edge_pec(mesh: Mesh, pec_tag: int) -> Set[int]:
    sel = mesh.tri_phys == pec_tag
    return set(int(e) for e in mesh.tri_edges[sel].reshape(-1))


def pec_mask(mesh: Mesh, pec: Set[int]) -> np.ndarray:
    m = mesh.num_edges
    mask = np.zeros(m, dtype=bool)
    if pec:
        idx = np.fromiter((e for e in pec if 0 <= e < m), dtype=np.int64)
        mask[idx] = True
    return mask


# --------------------------------------------------------------------------------------
# Element matrices  (src/edge_basis.cpp:14-130)
# --------------------------------------------------------------------------------------
def gradients_and_volume(X: np.ndarray):
    """X: [...,4,3] -> g [...,4,3], V [...]   (src/edge_basis.cpp:14-26)."""
    X = np.asarray(X, dtype=np.float64)
    # columns of B (B.col(k) = v[k]-v[3]); Eigen's fixed-size 3x3 inverse is cofactors * (1/det),
    # i.e. rows of B^-1 are cross products of the columns over det -- restated literally so the
    # oracle rounds like the reference rather than like LAPACK's pivoted LU.
    b0 = X[..., 0, :] - X[..., 3, :]
    b1 = X[..., 1, :] - X[..., 3, :]
    b2 = X[..., 2, :] - X[..., 3, :]
    c12 = np.cross(b1, b2)
    c20 = np.cross(b2, b0)
    c01 = np.cross(b0, b1)
    det = np.einsum("...k,...k->...", b0, c12)
    invdet = 1.0 / det
    g0 = c12 * invdet[..., None]
    g1 = c20 * invdet[..., None]
    g2 = c01 * invdet[..., None]
    g3 = -g0 - g1 - g2
    g = np.stack([g0, g1, g2, g3], axis=-2)
    V = np.abs(det) / 6.0
    return g, V


def whitney_curl_curl_matrix(X: np.ndarray) -> np.ndarray:
    """K(i,j) = V * c_i . c_j,  c_i = 2 grad(l_a) x grad(l_b)  (src/edge_basis.cpp:48-63)."""
    g, V = gradients_and_volume(X)
    ga = g[..., TET_PAIRS[:, 0], :]
    gb = g[..., TET_PAIRS[:, 1], :]
    curls = 2.0 * np.cross(ga, gb)
    K = np.einsum("...ik,...jk->...ij", curls, curls) * V[..., None, None]
    return K


def whitney_mass_matrix(X: np.ndarray) -> np.ndarray:
    """src/edge_basis.cpp:66-86 with lambda_int (src/edge_basis.cpp:28-30)."""
    g, V = gradients_and_volume(X)
    gg = np.einsum("...ak,...bk->...ab", g, g)  # g_a . g_b
    I = np.where(np.eye(4, dtype=bool), 1.0 / 10.0, 1.0 / 20.0)
    a = TET_PAIRS[:, 0][:, None]
    b = TET_PAIRS[:, 1][:, None]
    c = TET_PAIRS[:, 0][None, :]
    d = TET_PAIRS[:, 1][None, :]
    M = (
        gg[..., b, d] * I[a, c]
        - gg[..., b, c] * I[a, d]
        - gg[..., a, d] * I[b, c]
        + gg[..., a, c] * I[b, d]
    ) * V[..., None, None]
    return M


def triangle_whitney_mass_matrix(v: np.ndarray) -> np.ndarray:
    """3x3 surface edge mass on flat triangles; v [...,3,3]  (src/edge_basis.cpp:89-130)."""
    v = np.asarray(v, dtype=np.float64)
    normal = np.cross(v[..., 1, :] - v[..., 0, :], v[..., 2, :] - v[..., 0, :])
    area2 = np.linalg.norm(normal, axis=-1)
    safe = np.where(area2 < 1e-30, 1.0, area2)
    n_hat = normal / safe[..., None]
    area = area2 / 2.0
    g0 = np.cross(n_hat, v[..., 2, :] - v[..., 1, :]) / safe[..., None]
    g1 = np.cross(n_hat, v[..., 0, :] - v[..., 2, :]) / safe[..., None]
    g2 = np.cross(n_hat, v[..., 1, :] - v[..., 0, :]) / safe[..., None]
    g = np.stack([g0, g1, g2], axis=-2)
    gg = np.einsum("...ak,...bk->...ab", g, g)
    I = np.where(np.eye(3, dtype=bool), 1.0 / 6.0, 1.0 / 12.0)
    a = TRI_PAIRS[:, 0][:, None]
    b = TRI_PAIRS[:, 1][:, None]
    c = TRI_PAIRS[:, 0][None, :]
    d = TRI_PAIRS[:, 1][None, :]
    M = (
        gg[..., b, d] * I[a, c]
        - gg[..., b, c] * I[a, d]
        - gg[..., a, d] * I[b, c]
        + gg[..., a, c] * I[b, d]
    ) * area[..., None, None]
    M = np.where((area2 < 1e-30)[..., None, None], 0.0, M)
    return M


def triangle_mass_quadrature(v: np.ndarray) -> np.ndarray:
    """Independent check of the closed form: degree-2-exact 3-point edge-midpoint rule
    applied to N_e.N_f (quadratic), cf. tests/test_triangle_mass_matrix.cpp:74."""
    v = np.asarray(v, dtype=np.float64)
    normal = np.cross(v[1] - v[0], v[2] - v[0])
    area2 = np.linalg.norm(normal)
    n_hat = normal / area2
    g = [np.cross(n_hat, v[2] - v[1]) / area2, np.cross(n_hat, v[0] - v[2]) / area2, np.cross(n_hat, v[1] - v[0]) / area2]
    pts = [(0.5, 0.5, 0.0), (0.0, 0.5, 0.5), (0.5, 0.0, 0.5)]
    M = np.zeros((3, 3))
    for lam in pts:
        N = [lam[a] * g[b] - lam[b] * g[a] for a, b in TRI_PAIRS]
        for i in range(3):
            for j in range(3):
                M[i, j] += np.dot(N[i], N[j]) * (area2 / 2.0) / 3.0
    return M


# --------------------------------------------------------------------------------------
# Dispersive materials  (include/edgefem/materials/dispersive.hpp)
# --------------------------------------------------------------------------------------
class DispersiveMaterial:
    def eval_eps(self, omega: float) -> complex:  # pragma: no cover
        raise NotImplementedError

    def eval_mu(self, omega: float) -> complex:  # dispersive.hpp:28-31
        return 1.0 + 0.0j


class DebyeMaterial(DispersiveMaterial):
    def __init__(self, eps_static: float, eps_inf: float, tau: float):  # dispersive.hpp:55-60
        if tau <= 0.0:
            raise ValueError("DebyeMaterial: tau must be positive")
        self.eps_s, self.eps_inf, self.tau = eps_static, eps_inf, tau

    def eval_eps(self, omega):  # dispersive.hpp:62-66
        return self.eps_inf + (self.eps_s - self.eps_inf) / complex(1.0, omega * self.tau)


class LorentzMaterial(DispersiveMaterial):
    def __init__(self, eps_inf: float = 1.0):
        self.eps_inf = eps_inf
        self.poles: List[Tuple[float, float, float]] = []

    def add_pole(self, delta_eps, omega0, gamma):  # dispersive.hpp:116-125
        if omega0 <= 0.0:
            raise ValueError("LorentzMaterial: omega0 must be positive")
        if gamma < 0.0:
            raise ValueError("LorentzMaterial: gamma must be non-negative")
        self.poles.append((delta_eps, omega0, gamma))

    def eval_eps(self, omega):  # dispersive.hpp:127-138
        eps = complex(self.eps_inf, 0.0)
        w2 = omega * omega
        for de, w0, g in self.poles:
            w02 = w0 * w0
            eps += de * w02 / complex(w02 - w2, g * omega)
        return eps


class DrudeMaterial(DispersiveMaterial):
    def __init__(self, omega_p: float, gamma: float):  # dispersive.hpp:172-180
        if omega_p <= 0.0:
            raise ValueError("DrudeMaterial: omega_p must be positive")
        if gamma < 0.0:
            raise ValueError("DrudeMaterial: gamma must be non-negative")
        self.omega_p, self.gamma = omega_p, gamma

    def eval_eps(self, omega):  # dispersive.hpp:182-194
        if omega == 0.0:
            return complex(-1e30, 0.0)
        return 1.0 - (self.omega_p * self.omega_p) / complex(omega * omega, self.gamma * omega)


class DrudeLorentzMaterial(DispersiveMaterial):
    def __init__(self, eps_inf: float, omega_p: float, gamma_d: float):  # dispersive.hpp:224-235
        if omega_p <= 0.0:
            raise ValueError("DrudeLorentzMaterial: omega_p must be positive")
        if gamma_d < 0.0:
            raise ValueError("DrudeLorentzMaterial: gamma_d must be non-negative")
        self.eps_inf, self.omega_p, self.gamma_d = eps_inf, omega_p, gamma_d
        self.poles: List[Tuple[float, float, float]] = []

    def add_lorentz_pole(self, delta_eps, omega0, gamma):  # dispersive.hpp:238-249
        if omega0 <= 0.0:
            raise ValueError("DrudeLorentzMaterial: omega0 must be positive")
        if gamma < 0.0:
            raise ValueError("DrudeLorentzMaterial: gamma must be non-negative")
        self.poles.append((delta_eps, omega0, gamma))

    def eval_eps(self, omega):  # dispersive.hpp:251-273
        eps = complex(self.eps_inf, 0.0)
        if omega != 0.0:
            eps -= (self.omega_p * self.omega_p) / complex(omega * omega, self.gamma_d * omega)
        else:
            eps = complex(-1e30, 0.0)
        w2 = omega * omega
        for de, w0, g in self.poles:
            w02 = w0 * w0
            eps += de * w02 / complex(w02 - w2, g * omega)
        return eps


# --------------------------------------------------------------------------------------
# MaxwellParams / ports / periodic data  (include/edgefem/maxwell.hpp:24-149)
# --------------------------------------------------------------------------------------
@dataclass
class PMLRegionSpec:  # maxwell.hpp:24-28
    sigma_max: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    thickness: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    grading_order: float = 3.0


PORT_ABC_NONE, PORT_ABC_BETA, PORT_ABC_BETA_NORM, PORT_ABC_IMPEDANCE_MATCH, PORT_ABC_MODAL_ADMITTANCE = range(5)


@dataclass
class MaxwellParams:  # maxwell.hpp:46-143
    omega: float = 0.0
    eps_r: complex = 1.0
    mu_r: complex = 1.0
    eps_r_regions: Dict[int, complex] = field(default_factory=dict)
    mu_r_regions: Dict[int, complex] = field(default_factory=dict)
    eps_models: Dict[int, DispersiveMaterial] = field(default_factory=dict)
    mu_models: Dict[int, DispersiveMaterial] = field(default_factory=dict)
    pml_sigma: float = 0.0
    pml_regions: Set[int] = field(default_factory=set)
    pml_tensor_regions: Dict[int, PMLRegionSpec] = field(default_factory=dict)
    enforce_pml_heuristics: bool = True
    use_abc: bool = False
    abc_surface_tags: Set[int] = field(default_factory=set)
    use_port_abc: bool = False
    port_abc_type: int = PORT_ABC_BETA
    port_weight_scale: float = 1.0
    port_abc_scale: float = 1.0
    use_eigenmode_excitation: bool = False

    def get_eps_r(self, tag: int, omega: Optional[float] = None) -> complex:  # maxwell.hpp:100-122
        if omega is not None:
            mdl = self.eps_models.get(tag)
            if mdl is not None:
                return complex(mdl.eval_eps(omega))
        return complex(self.eps_r_regions.get(tag, self.eps_r))

    def get_mu_r(self, tag: int, omega: Optional[float] = None) -> complex:  # maxwell.hpp:124-142
        if omega is not None:
            mdl = self.mu_models.get(tag)
            if mdl is not None:
                return complex(mdl.eval_mu(omega))
        return complex(self.mu_r_regions.get(tag, self.mu_r))


@dataclass
class PortMode:  # ports/port_eigensolve.hpp:19-29
    pol: int = 0  # 0 TE, 1 TM
    fc: float = 0.0
    kc: float = 0.0
    omega: float = 0.0
    eps: complex = 0.0
    mu: complex = 0.0
    beta: complex = 0.0
    Z0: complex = 0.0


@dataclass
class WavePort:  # ports/wave_port.hpp:21-26
    surface_tag: int = 0
    mode: PortMode = field(default_factory=PortMode)
    edges: List[int] = field(default_factory=list)
    weights: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.complex128))


@dataclass
class PeriodicPair:  # periodic.hpp:13-19
    master_edge: int
    slave_edge: int
    master_orient: int
    slave_orient: int


@dataclass
class PeriodicBC:  # periodic.hpp:22-26
    pairs: List[PeriodicPair] = field(default_factory=list)
    period_vector: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    phase_shift: complex = 1.0 + 0.0j


# --------------------------------------------------------------------------------------
# Sparse helper: triplets -> CSR with duplicates summed and EXPLICIT ZEROS KEPT
# (Eigen setFromTriplets / coeffRef semantics)
# --------------------------------------------------------------------------------------
def triplets_to_csr(rows, cols, vals, m: int) -> sp.csr_matrix:
    A = sp.coo_matrix((np.asarray(vals), (np.asarray(rows), np.asarray(cols))), shape=(m, m)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


class _Trip:
    def __init__(self):
        self.r: List[np.ndarray] = []
        self.c: List[np.ndarray] = []
        self.v: List[np.ndarray] = []

    def add(self, r, c, v):
        r = np.asarray(r, dtype=np.int64).reshape(-1)
        c = np.asarray(c, dtype=np.int64).reshape(-1)
        v = np.asarray(v, dtype=np.complex128).reshape(-1)
        if v.size == 1 and r.size > 1:
            v = np.full(r.size, v[0], dtype=np.complex128)
        self.r.append(r)
        self.c.append(c)
        self.v.append(v)

    def csr(self, m: int) -> sp.csr_matrix:
        if not self.r:
            return sp.csr_matrix((m, m), dtype=np.complex128)
        return triplets_to_csr(np.concatenate(self.r), np.concatenate(self.c), np.concatenate(self.v), m)


# --------------------------------------------------------------------------------------
# assemble_maxwell  (src/assemble_maxwell.cpp:46-350)
# --------------------------------------------------------------------------------------
def pml_stretch(mesh: Mesh, p: MaxwellParams, X: np.ndarray) -> np.ndarray:
    """Per-tet scalar stretch s (src/assemble_maxwell.cpp:58-89,121-172)."""
    nt = mesh.tet_conn.shape[0]
    s = np.ones(nt, dtype=np.complex128)
    if p.omega == 0.0:
        return s
    for tag, spec in p.pml_tensor_regions.items():
        sel = np.nonzero(mesh.tet_phys == tag)[0]
        if sel.size == 0:
            continue
        pts = X[sel].reshape(-1, 3)
        bmin = pts.min(axis=0)
        bmax = pts.max(axis=0)
        cen = X[sel].sum(axis=1) / 4.0
        st = np.ones((sel.size, 3), dtype=np.complex128)
        kmin = 1e-3
        for ax in range(3):
            smax = spec.sigma_max[ax]
            th = spec.thickness[ax]
            if smax <= 0.0 or th <= 0.0:
                continue
            dmin = cen[:, ax] - bmin[ax]
            dmax = bmax[ax] - cen[:, ax]
            md = np.minimum(dmin, dmax)
            xi = np.where(md < th, 1.0 - md / th, 0.0)
            if p.enforce_pml_heuristics:
                xi = np.where(xi > 0.0, np.maximum(xi, kmin), xi)
            sig = smax * np.power(xi, spec.grading_order)
            if p.enforce_pml_heuristics and smax > 0.0:
                sig = np.where(sig < smax * kmin, smax * kmin, sig)
            st[:, ax] = 1.0 + 1j * sig / p.omega
        s[sel] = (st[:, 0] + st[:, 1] + st[:, 2]) / 3.0
    for tag in p.pml_regions:
        if tag in p.pml_tensor_regions:
            continue
        sel = mesh.tet_phys == tag
        s[sel] = complex(1.0, p.pml_sigma / p.omega)
    return s


def pml_diagnostics(p: MaxwellParams):
    """src/assemble_maxwell.cpp:91-112 -> list of (tag, sigma_max, thickness, reflection_est)."""
    out = []
    w = abs(p.omega)
    for tag, spec in p.pml_tensor_regions.items():
        refl = [1.0, 1.0, 1.0]
        for ax in range(3):
            smax, th = spec.sigma_max[ax], spec.thickness[ax]
            if smax <= 0.0 or th <= 0.0 or w <= 0.0:
                continue
            refl[ax] = math.exp(-2.0 * (smax * th / (spec.grading_order + 1.0)) / w)
        out.append((tag, tuple(spec.sigma_max), tuple(spec.thickness), tuple(refl)))
    return out


def material_tables(mesh: Mesh, p: MaxwellParams, dispersive: bool = True):
    tags = np.unique(mesh.tet_phys)
    eps = np.empty(mesh.tet_phys.shape[0], dtype=np.complex128)
    mu = np.empty_like(eps)
    for t in tags:
        sel = mesh.tet_phys == t
        if dispersive:
            eps[sel] = p.get_eps_r(int(t), p.omega)
            mu[sel] = p.get_mu_r(int(t), p.omega)
        else:
            eps[sel] = p.get_eps_r(int(t))
            mu[sel] = p.get_mu_r(int(t))
    return eps, mu


def volume_element_values(mesh: Mesh, p: MaxwellParams) -> np.ndarray:
    """[t,6,6] complex signed element matrices (src/assemble_maxwell.cpp:114-204)."""
    X = mesh.tet_xyz()
    s = pml_stretch(mesh, p, X)
    K = whitney_curl_curl_matrix(X).astype(np.complex128) / s[:, None, None]
    M = whitney_mass_matrix(X).astype(np.complex128) * s[:, None, None]
    k0 = p.omega / C0
    k0_sq = k0 * k0
    eps, mu = material_tables(mesh, p, dispersive=True)
    val = K / mu[:, None, None] - (k0_sq * eps)[:, None, None] * M
    sg = mesh.tet_orient.astype(np.float64)
    val = val * (sg[:, :, None] * sg[:, None, :])
    return val


def volume_abs_scale(mesh: Mesh, p: MaxwellParams) -> sp.csr_matrix:
    """sum over tets of |element contribution| per entry: the natural scale for comparing two
    floating-point summations of the same triplets (an entry whose contributions cancel cannot
    be reproduced to 1e-12 of its own magnitude by ANY summation order, Eigen's included)."""
    val = np.abs(volume_element_values(mesh, p))
    gi = np.repeat(mesh.tet_edges[:, :, None], 6, axis=2).reshape(-1)
    gj = np.repeat(mesh.tet_edges[:, None, :], 6, axis=1).reshape(-1)
    return triplets_to_csr(gi, gj, val.reshape(-1).astype(np.complex128), mesh.num_edges)


def volume_operand_scale(mesh: Mesh, p: MaxwellParams) -> sp.csr_matrix:
    """Per entry, the magnitude of the OPERANDS of the element arithmetic, summed over tets:
    V |c_i| |c_j| / |mu s| for the curl-curl dot product of src/edge_basis.cpp:59-61 and
    k0^2 |eps s| V (|g_b||g_d| I_ac + |g_b||g_c| I_ad + |g_a||g_d| I_bc + |g_a||g_c| I_bd) for the four mass terms of :76-82.
    Any correct fp64 evaluation (Eigen's included) is accurate to a few ulp of THIS scale: a dot product of
    near-orthogonal curls, or mass terms that cancel, lose relative accuracy in the result but not against their
    operands.  Test-side yardstick of the "1e-12 relative (fp64)" bar of north_star."""
    X = mesh.tet_xyz()
    g, V = gradients_and_volume(X)
    s = np.abs(pml_stretch(mesh, p, X))
    ga = g[..., TET_PAIRS[:, 0], :]
    gb = g[..., TET_PAIRS[:, 1], :]
    cn = np.linalg.norm(2.0 * np.cross(ga, gb), axis=-1)          # [t,6]
    gn = np.linalg.norm(g, axis=-1)                                # [t,4]
    I = np.where(np.eye(4, dtype=bool), 1.0 / 10.0, 1.0 / 20.0)
    a = TET_PAIRS[:, 0][:, None]
    b = TET_PAIRS[:, 1][:, None]
    c = TET_PAIRS[:, 0][None, :]
    d = TET_PAIRS[:, 1][None, :]
    Mb = (gn[:, b] * gn[:, d] * I[a, c] + gn[:, b] * gn[:, c] * I[a, d] + gn[:, a] * gn[:, d] * I[b, c] + gn[:, a] * gn[:, c] * I[b, d]) * V[:, None, None]
    Kb = cn[:, :, None] * cn[:, None, :] * V[:, None, None]
    k0 = p.omega / C0
    eps, mu = material_tables(mesh, p, dispersive=True)
    val = Kb / (np.abs(mu) * s)[:, None, None] + (k0 * k0 * np.abs(eps) * s)[:, None, None] * Mb
    gi = np.repeat(mesh.tet_edges[:, :, None], 6, axis=2).reshape(-1)
    gj = np.repeat(mesh.tet_edges[:, None, :], 6, axis=1).reshape(-1)
    return triplets_to_csr(gi, gj, val.reshape(-1).astype(np.complex128), mesh.num_edges)


def _abc_edges(mesh: Mesh, p: MaxwellParams, pec: Set[int]) -> np.ndarray:
    """src/assemble_maxwell.cpp:249-260 (order irrelevant: diagonal adds)."""
    if p.abc_surface_tags:
        sel = np.isin(mesh.tri_phys, list(p.abc_surface_tags))
    else:
        sel = np.ones(mesh.tri_phys.shape[0], dtype=bool)
    e = np.unique(mesh.tri_edges[sel].reshape(-1))
    if pec:
        e = e[~np.isin(e, list(pec))]
    return e


def _port_abc_coeff(p: MaxwellParams, port: WavePort) -> Optional[complex]:
    """src/assemble_maxwell.cpp:274-309."""
    if port.mode.Z0 == 0:
        return None
    k0 = p.omega / C0
    kc = port.mode.kc
    beta_sq = k0 * k0 - kc * kc
    if not beta_sq > 0:
        return None
    beta = math.sqrt(beta_sq)
    z0r = complex(port.mode.Z0).real
    t = p.port_abc_type
    if t == PORT_ABC_BETA:
        return complex(0.0, beta)
    if t == PORT_ABC_BETA_NORM:
        return complex(0.0, beta / k0)
    if t == PORT_ABC_IMPEDANCE_MATCH:
        return complex(0.0, beta * math.sqrt(z0r / ETA0))
    if t == PORT_ABC_MODAL_ADMITTANCE:
        return complex(0.0, p.omega * EPS0 / z0r)
    return 0.0 + 0.0j


def assemble_maxwell(mesh: Mesh, p: MaxwellParams, pec: Set[int], ports: Sequence[WavePort] = (), active_port_idx: int = -1):
    """Returns (A csr complex128 with explicit zeros kept, b complex128[m])."""
    m = mesh.num_edges
    trip = _Trip()
    val = volume_element_values(mesh, p)
    gi = np.repeat(mesh.tet_edges[:, :, None], 6, axis=2)
    gj = np.repeat(mesh.tet_edges[:, None, :], 6, axis=1)
    trip.add(gi, gj, val)
    b = np.zeros(m, dtype=np.complex128)
    pmask = pec_mask(mesh, pec)

    # ports: dense block + source (assemble_maxwell.cpp:212-242)
    for i, port in enumerate(ports):
        w = np.asarray(port.weights, dtype=np.complex128)
        if w.size != len(port.edges):
            continue
        if port.mode.Z0 == 0:
            continue
        e = np.asarray(port.edges, dtype=np.int64)
        free = ~pmask[e]
        ef, wf = e[free], w[free]
        blk = np.outer(wf, np.conj(wf)) / port.mode.Z0
        trip.add(np.repeat(ef[:, None], ef.size, 1), np.repeat(ef[None, :], ef.size, 0), blk)
        if i == active_port_idx:
            scale = 2.0 / np.sqrt(complex(port.mode.Z0))
            np.add.at(b, ef, scale * wf)

    if p.use_abc:  # assemble_maxwell.cpp:244-265
        e = _abc_edges(mesh, p, pec)
        trip.add(e, e, np.full(e.size, complex(0.0, p.omega / C0)))

    if p.use_port_abc and p.port_abc_type != PORT_ABC_NONE:  # assemble_maxwell.cpp:274-320
        for port in ports:
            coeff = _port_abc_coeff(p, port)
            if coeff is None:
                continue
            e = np.asarray(port.edges, dtype=np.int64)
            e = e[~pmask[e]]
            trip.add(e, e, np.full(e.size, coeff))

    # Dirichlet diag insertion (coeffRef(e,e)=1 creates the entry if absent) :344-347
    pe = np.nonzero(pmask)[0]
    trip.add(pe, pe, np.zeros(pe.size))
    A = trip.csr(m)
    apply_dirichlet(A, pmask)
    b[pmask] = 0.0
    return A, b


def apply_dirichlet(A: sp.csr_matrix, pmask: np.ndarray, diag_value: complex = 1.0) -> None:
    """Zero PEC rows+cols keeping the entries, diagonal=1 (src/assemble_maxwell.cpp:322-347)."""
    m = A.shape[0]
    rows = np.repeat(np.arange(m), np.diff(A.indptr))
    kill = pmask[rows] | pmask[A.indices]
    A.data[kill] = 0.0
    diag = (rows == A.indices) & pmask[rows]
    A.data[diag] = diag_value


# --------------------------------------------------------------------------------------
# Port builders  (src/ports/*.cpp)
# --------------------------------------------------------------------------------------
def solve_te10_mode(a: float, b: float, freq: float) -> PortMode:
    """src/ports/port_eigensolve.cpp:45-63 (eta0 literal at :15)."""
    mode = PortMode()
    mode.pol = 0
    mode.fc = C0 / (2.0 * a)
    k = 2.0 * math.pi * freq / C0
    kc = math.pi / a
    mode.kc = kc
    mode.omega = 2.0 * math.pi * freq
    mode.mu = MU0
    mode.eps = EPS0
    beta = math.sqrt(max(0.0, k * k - kc * kc))
    mode.beta = beta
    mode.Z0 = 376.730313668 * k / beta if beta > 0 else 0.0
    return mode


def assemble_port_surface_mass(mesh: Mesh, surface_tag: int, pec: Set[int]) -> sp.csr_matrix:
    """Real m x m M_s; PEC rows/cols omitted (src/ports/wave_port.cpp:549-585)."""
    m = mesh.num_edges
    sel = np.nonzero(mesh.tri_phys == surface_tag)[0]
    if sel.size == 0:
        return sp.csr_matrix((m, m), dtype=np.float64)
    v = mesh.tri_xyz()[sel]
    Ml = triangle_whitney_mass_matrix(v)
    o = mesh.tri_orient[sel].astype(np.float64)
    vals = Ml * (o[:, :, None] * o[:, None, :])
    e = mesh.tri_edges[sel]
    gi = np.repeat(e[:, :, None], 3, 2).reshape(-1)
    gj = np.repeat(e[:, None, :], 3, 1).reshape(-1)
    vals = vals.reshape(-1)
    pmask = pec_mask(mesh, pec)
    keep = ~(pmask[gi] | pmask[gj])
    M = sp.coo_matrix((vals[keep], (gi[keep], gj[keep])), shape=(m, m)).tocsr()
    M.sum_duplicates()
    M.sort_indices()
    return M


def port_2d_matrices(mesh: Mesh, surface_tag: int, pec: Set[int]):
    """Dense K_s, M_s on free port edges in first-seen order (src/ports/wave_port.cpp:416-485)."""
    port_edges: List[int] = []
    local: Dict[int, int] = {}
    sel = np.nonzero(mesh.tri_phys == surface_tag)[0]
    for t in sel:
        for le in range(3):
            ge = int(mesh.tri_edges[t, le])
            if ge in pec or ge in local:
                continue
            local[ge] = len(port_edges)
            port_edges.append(ge)
    n = len(port_edges)
    K = np.zeros((n, n))
    M = np.zeros((n, n))
    txyz = mesh.tri_xyz()
    for t in sel:
        v = txyz[t]
        normal = np.cross(v[1] - v[0], v[2] - v[0])
        area2 = np.linalg.norm(normal)
        if area2 < 1e-30:
            continue
        n_hat = normal / area2
        area = area2 / 2.0
        g = [np.cross(n_hat, v[2] - v[1]) / area2, np.cross(n_hat, v[0] - v[2]) / area2, np.cross(n_hat, v[1] - v[0]) / area2]
        Ml = triangle_whitney_mass_matrix(v)
        curls = [2.0 * np.dot(np.cross(g[a], g[b]), n_hat) for a, b in TRI_PAIRS]
        for i in range(3):
            li = local.get(int(mesh.tri_edges[t, i]))
            if li is None:
                continue
            for j in range(3):
                lj = local.get(int(mesh.tri_edges[t, j]))
                if lj is None:
                    continue
                sign = float(mesh.tri_orient[t, i] * mesh.tri_orient[t, j])
                K[li, lj] += sign * curls[i] * curls[j] * area
                M[li, lj] += sign * Ml[i, j]
    return port_edges, K, M


def solve_port_mode_2d(mesh: Mesh, surface_tag: int, pec: Set[int], target_kc_sq: float):
    """Returns (v_full real[m], kc_sq) (src/ports/wave_port.cpp:410-514).
    Eigenvector sign/scale are solver-defined (v^T M v = 1 here, like Eigen's GSAES)."""
    m = mesh.num_edges
    port_edges, K, M = port_2d_matrices(mesh, surface_tag, pec)
    if not port_edges:
        return np.zeros(m), 0.0
    ev, V = scipy.linalg.eigh(K, M)
    kc_min = 1e-6 * max(target_kc_sq, 1.0)
    best, best_dist = -1, float("inf")
    for i in range(len(ev)):
        if ev[i] < kc_min:
            continue
        d = abs(ev[i] - target_kc_sq)
        if d < best_dist:
            best_dist, best = d, i
    if best < 0:
        return np.zeros(m), 0.0
    v_full = np.zeros(m)
    v_full[np.asarray(port_edges)] = V[:, best]
    return v_full, float(ev[best])


def build_wave_port_2d(mesh: Mesh, surface_tag: int, mode: PortMode, pec: Set[int], target_kc_sq: float) -> WavePort:
    """src/ports/wave_port.cpp:516-546."""
    port = WavePort(surface_tag=surface_tag, mode=PortMode(**vars(mode)))
    v, kc_sq = solve_port_mode_2d(mesh, surface_tag, pec, target_kc_sq)
    if kc_sq > 0.0:
        port.mode.kc = math.sqrt(kc_sq)
    sel = mesh.tri_phys == surface_tag
    port.edges = sorted(set(int(e) for e in mesh.tri_edges[sel].reshape(-1)))
    port.weights = v[np.asarray(port.edges, dtype=np.int64)].astype(np.complex128)
    return port


def build_lumped_port(mesh: Mesh, surface_tag: int, z0: float = 50.0, e_direction=(0.0, 0.0, 1.0), surface_integral: bool = True) -> WavePort:
    """src/ports/lumped_port.cpp:21-160.  Edge order: sorted (the reference iterates an
    unordered_set -> implementation-defined; results are order-independent up to rounding)."""
    sel = np.nonzero(mesh.tri_phys == surface_tag)[0]
    edge_set = sorted(set(int(e) for e in mesh.tri_edges[sel].reshape(-1)))
    if not edge_set:
        raise RuntimeError("build_lumped_port: no triangles found with surface_tag = %d" % surface_tag)
    idx = {e: i for i, e in enumerate(edge_set)}
    e_dir = np.asarray(e_direction, dtype=np.float64)
    e_dir = e_dir / np.linalg.norm(e_dir)
    w = np.zeros(len(edge_set), dtype=np.complex128)
    if surface_integral:
        txyz = mesh.tri_xyz()
        for t in sel:
            v = txyz[t]
            normal = np.cross(v[1] - v[0], v[2] - v[0])
            area2 = np.linalg.norm(normal)
            if area2 < 1e-30:
                continue
            n_hat = normal / area2
            area = area2 / 2.0
            g = [np.cross(n_hat, v[2] - v[1]) / area2, np.cross(n_hat, v[0] - v[2]) / area2, np.cross(n_hat, v[1] - v[0]) / area2]
            for le in range(3):
                li, lj = TRI_PAIRS[le]
                integral = (area / 3.0) * np.dot(g[lj] - g[li], e_dir)
                w[idx[int(mesh.tri_edges[t, le])]] += mesh.tri_orient[t, le] * integral
    else:
        for i, e in enumerate(edge_set):
            n0, n1 = mesh.edges[e]
            p0 = mesh.xyz[mesh.node_idx_of(np.array([n0]))[0]]
            p1 = mesh.xyz[mesh.node_idx_of(np.array([n1]))[0]]
            w[i] = np.dot(p1 - p0, e_dir)
    nrm = np.linalg.norm(w)
    if nrm > 1e-15:
        w *= math.sqrt(z0) / nrm
    port = WavePort(surface_tag=surface_tag)
    port.edges = edge_set
    port.weights = w
    port.mode.Z0 = z0
    return port


# --------------------------------------------------------------------------------------
# Solve  (src/solver.cpp:11-33 -- SparseLU branch; SuperLU here)
# --------------------------------------------------------------------------------------
def solve_direct(A: sp.csr_matrix, b: np.ndarray):
    lu = spla.splu(A.tocsc())
    x = lu.solve(b)
    nb = np.linalg.norm(b)
    res = np.linalg.norm(A @ x - b) / nb if nb > 0 else 0.0
    return x, res


# --------------------------------------------------------------------------------------
# S-parameter drivers
# --------------------------------------------------------------------------------------
def _project(ports, pmask, x, i, j):
    """V_j = sum conj(w_jk) x(edge_jk) over non-PEC edges (src/assemble_maxwell.cpp:376-384)."""
    e = np.asarray(ports[j].edges, dtype=np.int64)
    w = np.asarray(ports[j].weights, dtype=np.complex128)
    free = ~pmask[e]
    return np.sum(np.conj(w[free]) * x[e[free]])


def calculate_sparams(mesh, p, pec, ports) -> np.ndarray:
    """src/assemble_maxwell.cpp:352-394."""
    n = len(ports)
    S = np.zeros((n, n), dtype=np.complex128)
    pmask = pec_mask(mesh, pec)
    for i in range(n):
        A, b = assemble_maxwell(mesh, p, pec, ports, i)
        x, _ = solve_direct(A, b)
        vinc = np.sqrt(complex(ports[i].mode.Z0))
        for j in range(n):
            vj = _project(ports, pmask, x, i, j)
            S[j, i] = (vj - vinc) / vinc if i == j else vj / vinc
    return S


def normalize_port_weights(mesh, p, pec, ports) -> None:
    """In-place (src/assemble_maxwell.cpp:396-494)."""
    if not ports:
        return
    A, _ = assemble_maxwell(mesh, p, pec, [], -1)
    pmask = pec_mask(mesh, pec)
    m = mesh.num_edges
    for port in ports:
        w = np.asarray(port.weights, dtype=np.complex128)
        if w.size != len(port.edges) or port.mode.Z0 == 0:
            continue
        e = np.asarray(port.edges, dtype=np.int64)
        free = ~pmask[e]
        rhs = np.zeros(m, dtype=np.complex128)
        rhs[e[free]] = w[free]
        y, _ = solve_direct(A, rhs)
        if np.linalg.norm(w) < 1e-15:
            continue
        wAw = abs(np.sum(np.conj(w[free]) * y[e[free]]))
        target = complex(port.mode.Z0).real
        if wAw > 1e-15 and target > 1e-15:
            port.weights = w * math.sqrt(target / wAw)


def port_betas(mesh, p, ports):
    """beta_i = sqrt(eps mu k0^2 - kc^2) with the material of the first tet (file order) having
    a face on the port; None if evanescent (src/assemble_maxwell.cpp:652-700)."""
    k0 = p.omega / C0
    faces = [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]
    betas = []
    for port in ports:
        sel = mesh.tri_phys == port.surface_tag
        pf = set(tuple(sorted(int(x) for x in c)) for c in mesh.tri_conn[sel])
        eps_mu = 1.0 + 0.0j
        if pf:
            keys = []
            for f in faces:
                keys.append(np.sort(mesh.tet_conn[:, list(f)], axis=1))
            found = None
            # first tet in file order with any face (faces checked in order) on the port
            pf_arr = np.array(sorted(pf), dtype=np.int64)
            pf_key = pf_arr[:, 0] * (1 << 42) + pf_arr[:, 1] * (1 << 21) + pf_arr[:, 2]
            hit = np.zeros(mesh.tet_conn.shape[0], dtype=bool)
            for kf in keys:
                kk = kf[:, 0] * (1 << 42) + kf[:, 1] * (1 << 21) + kf[:, 2]
                hit |= np.isin(kk, pf_key)
            nz = np.nonzero(hit)[0]
            if nz.size:
                found = int(nz[0])
            if found is not None:
                tag = int(mesh.tet_phys[found])
                eps_mu = p.get_eps_r(tag, p.omega) * p.get_mu_r(tag, p.omega)
        beta_sq = eps_mu * k0 * k0 - port.mode.kc * port.mode.kc
        if complex(beta_sq).real <= 0:
            return None
        betas.append(np.sqrt(complex(beta_sq)))
    return betas


def eigenmode_system(mesh, p, pec, ports):
    """Shared pieces of calculate_sparams_eigenmode: (A_base+sum j*scale*beta_i*M_s,i , port_mass, port_vecs, betas)."""
    m = mesh.num_edges
    betas = port_betas(mesh, p, ports)
    if betas is None:
        return None
    pmask = pec_mask(mesh, pec)
    port_mass = [assemble_port_surface_mass(mesh, port.surface_tag, pec) for port in ports]
    port_vecs = []
    for i, port in enumerate(ports):
        v = np.zeros(m, dtype=np.complex128)
        e = np.asarray(port.edges, dtype=np.int64)
        w = np.asarray(port.weights, dtype=np.complex128)
        free = ~pmask[e]
        v[e[free]] = w[free]
        nsq = np.real(np.vdot(v, port_mass[i] @ v))
        if nsq > 1e-30:
            v = v / math.sqrt(nsq)
        port_vecs.append(v)
    A0, _ = assemble_maxwell(mesh, p, pec, [], -1)
    trip = _Trip()
    coo = A0.tocoo()
    trip.add(coo.row, coo.col, coo.data)
    for i in range(len(ports)):
        coeff = complex(0.0, p.port_abc_scale) * betas[i]
        c = port_mass[i].tocoo()
        trip.add(c.row, c.col, coeff * c.data)
    A = trip.csr(m)
    return A, port_mass, port_vecs, betas


def calculate_sparams_eigenmode(mesh, p, pec, ports, return_fields: bool = False):
    """src/assemble_maxwell.cpp:637-787."""
    n = len(ports)
    S = np.zeros((n, n), dtype=np.complex128)
    if n == 0:
        return S
    sysm = eigenmode_system(mesh, p, pec, ports)
    if sysm is None:
        return S  # reference returns an uninitialised S after a warning (:687-698)
    A, port_mass, port_vecs, betas = sysm
    lu = spla.splu(A.tocsc())
    fields = []
    for a in range(n):
        abc_a = complex(0.0, p.port_abc_scale) * betas[a]
        b = 2.0 * abc_a * (port_mass[a] @ port_vecs[a])
        x = lu.solve(b)
        fields.append(x)
        for j in range(n):
            vj = np.vdot(port_vecs[j], port_mass[j] @ x)
            S[j, a] = vj - 1.0 if j == a else vj
    if return_fields:
        return S, fields
    return S


# --------------------------------------------------------------------------------------
# Periodic  (src/periodic.cpp:48-225, src/assemble_maxwell.cpp:496-635)
# --------------------------------------------------------------------------------------
def edge_centroids(mesh: Mesh, eidx: np.ndarray) -> np.ndarray:
    n0 = mesh.node_idx_of(mesh.edges[eidx, 0])
    n1 = mesh.node_idx_of(mesh.edges[eidx, 1])
    return 0.5 * (mesh.xyz[n0] + mesh.xyz[n1])


def build_periodic_pairs(mesh: Mesh, master_tag: int, slave_tag: int, period_vector, tolerance: float = 1e-9) -> PeriodicBC:
    """src/periodic.cpp:48-164.  Pair order: ascending master edge id (the reference iterates
    an unordered_set, i.e. implementation-defined order)."""
    pv = np.asarray(period_vector, dtype=np.float64)
    pbc = PeriodicBC(period_vector=tuple(pv))
    msel = mesh.tri_phys == master_tag
    ssel = mesh.tri_phys == slave_tag
    medges = np.unique(mesh.tri_edges[msel].reshape(-1))
    sedges = np.unique(mesh.tri_edges[ssel].reshape(-1))
    if medges.size == 0:
        raise RuntimeError("No edges found on master surface with tag %d" % master_tag)
    if sedges.size == 0:
        raise RuntimeError("No edges found on slave surface with tag %d" % slave_tag)
    sc = edge_centroids(mesh, sedges)
    mc = edge_centroids(mesh, medges) + pv
    matched: Set[int] = set()

    def orient_on(sel, e):
        o = 0
        for t in np.nonzero(sel)[0]:
            for i in range(3):
                if mesh.tri_edges[t, i] == e:
                    o = int(mesh.tri_orient[t, i])
                    break  # edge_orientation_on_tri returns the first hit per tri
        return o if o != 0 else 1

    # precompute "last matching tri" orientation per edge (periodic.cpp:135-146)
    def orient_table(sel):
        tab: Dict[int, int] = {}
        for t in np.nonzero(sel)[0]:
            seen = set()
            for i in range(3):
                e = int(mesh.tri_edges[t, i])
                if e in seen:
                    continue
                seen.add(e)
                tab[e] = int(mesh.tri_orient[t, i])
        return tab

    mtab, stab = orient_table(msel), orient_table(ssel)
    for k, me in enumerate(medges):
        d = np.linalg.norm(sc - mc[k], axis=1)
        hits = np.nonzero(d < tolerance)[0]
        if hits.size == 0:
            raise RuntimeError("Could not find matching slave edge for master edge %d" % int(me))
        se = int(sedges[hits[0]])
        if se in matched:
            raise RuntimeError("Slave edge %d matched to multiple master edges" % se)
        matched.add(se)
        pbc.pairs.append(PeriodicPair(int(me), se, mtab.get(int(me), 1), stab.get(se, 1)))
    return pbc


def set_floquet_phase(pbc: PeriodicBC, kx: float, ky: float) -> None:
    """src/periodic.cpp:204-211."""
    arg = kx * pbc.period_vector[0] + ky * pbc.period_vector[1]
    pbc.phase_shift = complex(math.cos(arg), math.sin(arg))


def floquet_phase_from_angle(period_vector, theta: float, phi: float, k0: float) -> complex:
    """src/periodic.cpp:213-225."""
    k = (k0 * math.sin(theta) * math.cos(phi), k0 * math.sin(theta) * math.sin(phi), -k0 * math.cos(theta))
    arg = k[0] * period_vector[0] + k[1] * period_vector[1] + k[2] * period_vector[2]
    return complex(math.cos(arg), math.sin(arg))


def assemble_maxwell_periodic_dense(mesh, p, pec, pbc: PeriodicBC, ports=(), active_port_idx=-1):
    """Literal dense restatement of src/assemble_maxwell.cpp:496-574 (small meshes only).
    Returns (A csr with exact zeros dropped like sparseView(), b)."""
    A, b = assemble_maxwell(mesh, p, pec, ports, active_port_idx)
    if not pbc.pairs:
        return A, b
    Ad = A.toarray()
    bn = b.copy()
    phi = complex(pbc.phase_shift)
    for pr in pbc.pairs:
        orient = float(pr.master_orient * pr.slave_orient)
        ph = phi * orient
        phc = np.conj(phi) * orient
        if pr.master_edge in pec or pr.slave_edge in pec:
            continue
        Ad[pr.master_edge, :] += ph * Ad[pr.slave_edge, :]
        Ad[:, pr.master_edge] += phc * Ad[:, pr.slave_edge]
        bn[pr.master_edge] += ph * bn[pr.slave_edge]
    for pr in pbc.pairs:
        s = pr.slave_edge
        if s in pec:
            continue
        Ad[s, :] = 0.0
        Ad[:, s] = 0.0
        Ad[s, s] = 1.0
        bn[s] = 0.0
    As = sp.csr_matrix(Ad)
    As.eliminate_zeros()
    As.sort_indices()
    return As, bn


def assemble_maxwell_periodic(mesh, p, pec, pbc: PeriodicBC, ports=(), active_port_idx=-1):
    """Sparse equivalent of the dense procedure: A' = T A T^H with T = I + sum phase_k e_m e_s^T
    (valid when no edge is both master and slave and no edge is in two pairs -- true for a
    single build_periodic_pairs() result), then slave rows/cols -> identity, exact zeros dropped."""
    A, b = assemble_maxwell(mesh, p, pec, ports, active_port_idx)
    if not pbc.pairs:
        return A, b
    m = mesh.num_edges
    phi = complex(pbc.phase_shift)
    act = [pr for pr in pbc.pairs if not (pr.master_edge in pec or pr.slave_edge in pec)]
    masters = set(pr.master_edge for pr in act)
    slaves = set(pr.slave_edge for pr in act)
    if masters & slaves or len(masters) != len(act) or len(slaves) != len(act):
        return assemble_maxwell_periodic_dense(mesh, p, pec, pbc, ports, active_port_idx)
    rows = [pr.master_edge for pr in act]
    cols = [pr.slave_edge for pr in act]
    vals = [phi * float(pr.master_orient * pr.slave_orient) for pr in act]
    T = sp.identity(m, dtype=np.complex128, format="csr") + sp.csr_matrix((vals, (rows, cols)), shape=(m, m))
    Ap = (T @ A @ T.conj().T).tolil()
    bn = T @ b
    for pr in pbc.pairs:
        s = pr.slave_edge
        if s in pec:
            continue
        Ap[s, :] = 0.0
        Ap[:, s] = 0.0
        Ap[s, s] = 1.0
        bn[s] = 0.0
    Ap = Ap.tocsr()
    Ap.eliminate_zeros()
    Ap.sort_indices()
    return Ap, bn


def calculate_sparams_periodic(mesh, p, pec, pbc: PeriodicBC, ports, dense: bool = False):
    """src/assemble_maxwell.cpp:576-635."""
    n = len(ports)
    S = np.zeros((n, n), dtype=np.complex128)
    pmask = pec_mask(mesh, pec)
    asm = assemble_maxwell_periodic_dense if dense else assemble_maxwell_periodic
    for i in range(n):
        A, b = asm(mesh, p, pec, pbc, ports, i)
        x, _ = solve_direct(A, b)
        xf = x.copy()
        for pr in pbc.pairs:
            xf[pr.slave_edge] = pbc.phase_shift * float(pr.master_orient * pr.slave_orient) * xf[pr.master_edge]
        vinc = np.sqrt(complex(ports[i].mode.Z0))
        for j in range(n):
            vj = _project(ports, pmask, xf, i, j)
            S[j, i] = (vj - vinc) / vinc if i == j else vj / vinc
    return S


# --------------------------------------------------------------------------------------
# Sweep  (src/sweep.cpp:82-349)
# --------------------------------------------------------------------------------------
def assemble_maxwell_km(mesh, p, pec):
    """K <- sum (K_loc/mu) ss, M <- sum (eps M_loc) ss with STATIC materials; Dirichlet rows/cols
    zeroed in both, K(e,e)=1 (src/sweep.cpp:90-172)."""
    m = mesh.num_edges
    X = mesh.tet_xyz()
    eps, mu = material_tables(mesh, p, dispersive=False)
    sg = mesh.tet_orient.astype(np.float64)
    ss = sg[:, :, None] * sg[:, None, :]
    Kv = (whitney_curl_curl_matrix(X).astype(np.complex128) / mu[:, None, None]) * ss
    Mv = (eps[:, None, None] * whitney_mass_matrix(X).astype(np.complex128)) * ss
    gi = np.repeat(mesh.tet_edges[:, :, None], 6, axis=2).reshape(-1)
    gj = np.repeat(mesh.tet_edges[:, None, :], 6, axis=1).reshape(-1)
    pmask = pec_mask(mesh, pec)
    pe = np.nonzero(pmask)[0]
    K = triplets_to_csr(np.concatenate([gi, pe]), np.concatenate([gj, pe]), np.concatenate([Kv.reshape(-1), np.zeros(pe.size)]), m)
    M = triplets_to_csr(gi, gj, Mv.reshape(-1), m)
    apply_dirichlet(K, pmask, 1.0)
    rows = np.repeat(np.arange(m), np.diff(M.indptr))
    M.data[pmask[rows] | pmask[M.indices]] = 0.0
    return K, M


def frequency_sweep(mesh, p: MaxwellParams, pec, ports, freqs):
    """S[F,P,P] (src/sweep.cpp:174-349), direct solves."""
    n = len(ports)
    out = []
    if n == 0 or len(freqs) == 0:
        return np.zeros((0, n, n), dtype=np.complex128)
    has_pml = bool(p.pml_regions) or bool(p.pml_tensor_regions)
    has_disp = bool(p.eps_models) or bool(p.mu_models)
    pmask = pec_mask(mesh, pec)
    m = mesh.num_edges
    if not has_pml and not has_disp:
        K, M = assemble_maxwell_km(mesh, p, pec)
        trip_base = _Trip()
        for port in ports:
            w = np.asarray(port.weights, dtype=np.complex128)
            if w.size != len(port.edges) or port.mode.Z0 == 0:
                continue
            e = np.asarray(port.edges, dtype=np.int64)
            free = ~pmask[e]
            ef, wf = e[free], w[free]
            trip_base.add(np.repeat(ef[:, None], ef.size, 1), np.repeat(ef[None, :], ef.size, 0), np.outer(wf, np.conj(wf)) / port.mode.Z0)
        abc_e = _abc_edges(mesh, p, pec) if p.use_abc else np.zeros(0, dtype=np.int64)
        Kc, Mc = K.tocoo(), M.tocoo()
        for f in freqs:
            omega = 2.0 * math.pi * f
            k0 = omega / C0
            trip = _Trip()
            trip.add(Kc.row, Kc.col, Kc.data)
            trip.add(Mc.row, Mc.col, -(k0 * k0) * Mc.data)
            trip.r += trip_base.r
            trip.c += trip_base.c
            trip.v += trip_base.v
            if p.use_abc:
                trip.add(abc_e, abc_e, np.full(abc_e.size, complex(0.0, k0)))
            A = trip.csr(m)
            lu = spla.splu(A.tocsc())
            S = np.zeros((n, n), dtype=np.complex128)
            for a in range(n):
                b = np.zeros(m, dtype=np.complex128)
                e = np.asarray(ports[a].edges, dtype=np.int64)
                w = np.asarray(ports[a].weights, dtype=np.complex128)
                free = ~pmask[e]
                np.add.at(b, e[free], (2.0 / np.sqrt(complex(ports[a].mode.Z0))) * w[free])
                x = lu.solve(b)
                vinc = np.sqrt(complex(ports[a].mode.Z0))
                for j in range(n):
                    vj = _project(ports, pmask, x, a, j)
                    S[j, a] = (vj - vinc) / vinc if j == a else vj / vinc
            out.append(S)
    else:
        for f in freqs:
            pf = MaxwellParams(**{k: v for k, v in vars(p).items()})
            pf.omega = 2.0 * math.pi * f
            out.append(calculate_sparams(mesh, pf, pec, ports))
    return np.array(out)


# --------------------------------------------------------------------------------------
# Convenience: the WR-90 benchmark flow (tests/benchmark_wr90.cpp:32-95)
# --------------------------------------------------------------------------------------
def wr90_ports(mesh: Mesh, pec: Set[int], freq: float, a: float = 0.02286, b: float = 0.01016):
    kc_sq = (math.pi / a) ** 2
    m1 = solve_te10_mode(a, b, freq)
    m2 = solve_te10_mode(a, b, freq)
    return [build_wave_port_2d(mesh, 2, m1, pec, kc_sq), build_wave_port_2d(mesh, 3, m2, pec, kc_sq)]


def wr90_sparams(mesh: Mesh, pec: Set[int], freq: float, ports=None, port_abc_scale: float = 1.0) -> np.ndarray:
    if ports is None:
        ports = wr90_ports(mesh, pec, freq)
    p = MaxwellParams(omega=2.0 * math.pi * freq, port_abc_scale=port_abc_scale)
    return calculate_sparams_eigenmode(mesh, p, pec, ports)


# =====================================================================================================
# nodal port eigenmodes, port-face extraction, modal line-integral ports,
# Touchstone text.  Test infrastructure like the rest of this file.
# =====================================================================================================
@dataclass
class PortSurface:  # ports/wave_port.hpp:14-17 (PortSurfaceMesh): 2-D mesh of the port face + source tri indices
    node_ids: np.ndarray  # int64 [n]   first-seen order over the face's tris
    xy: np.ndarray  # float64 [n,2]
    tri_conn: np.ndarray  # int64 [k,3] node ids
    volume_tri_indices: np.ndarray  # int64 [k]
    boundary_lines: np.ndarray  # int64 [l,2] (n0<n1) edges seen once, phys 1
    phys: int = 0

    def idx(self, ids):
        lut = {int(n): i for i, n in enumerate(self.node_ids)}
        return np.vectorize(lut.__getitem__)(np.asarray(ids))


def extract_surface_mesh(mesh: Mesh, surface_tag: int) -> PortSurface:  # src/ports/wave_port.cpp:54-112
    sel = np.nonzero(mesh.tri_phys == surface_tag)[0]
    order, seen = [], set()
    for t in sel:
        for nid in mesh.tri_conn[t]:
            if int(nid) not in seen:
                seen.add(int(nid))
                order.append(int(nid))
    node_ids = np.array(order, dtype=np.int64)
    xy = mesh.xyz[mesh.node_idx_of(node_ids)][:, :2] if node_ids.size else np.zeros((0, 2))
    cnt: Dict[Tuple[int, int], int] = {}
    for t in sel:
        c = mesh.tri_conn[t]
        for e in range(3):
            a, b = int(c[e]), int(c[(e + 1) % 3])
            k = (min(a, b), max(a, b))
            cnt[k] = cnt.get(k, 0) + 1
    bl = np.array(sorted(k for k, v in cnt.items() if v == 1), dtype=np.int64).reshape(-1, 2)
    return PortSurface(node_ids, xy, mesh.tri_conn[sel].copy(), sel.astype(np.int64), bl, int(surface_tag) if sel.size else 0)


def _tri_shape(xy3: np.ndarray):
    """P1 gradients (rows = nodes, cols = d/dx, d/dy) = columns 1..2 of inv([[1,x,y]]) and |area| (port_eigensolve.cpp:139-146)."""
    C = np.column_stack([np.ones(3), xy3[:, 0], xy3[:, 1]])
    return np.linalg.inv(C)[1:3, :].T.copy(), 0.5 * abs(np.linalg.det(C))


def solve_port_eigens(surf: PortSurface, num_modes: int, omega: float, eps_r: complex, mu_r: complex, tm: bool = False):
    """src/ports/port_eigensolve.cpp:97-275.  Returns list of (PortMode, field[n]).  Sign: largest sample positive."""
    if surf.tri_conn.shape[0] == 0:
        raise RuntimeError("Port eigensolver requires a 2D mesh.")
    if omega == 0.0:
        raise RuntimeError("Port eigensolver requires non-zero frequency.")
    import scipy.linalg as sla

    nn = surf.node_ids.size
    pec = np.zeros(nn, dtype=bool)
    if tm and surf.boundary_lines.size:
        pec[surf.idx(surf.boundary_lines.reshape(-1))] = True
    dof = np.full(nn, -1)
    dof[~pec] = np.arange(int((~pec).sum()))
    nd = int((~pec).sum())
    A = np.zeros((nd, nd))
    B = np.zeros((nd, nd))
    tri_idx = surf.idx(surf.tri_conn)
    shapes = [_tri_shape(surf.xy[tri_idx[t]]) for t in range(tri_idx.shape[0])]
    for t in range(tri_idx.shape[0]):
        G, area = shapes[t]
        ke = area * G @ G.T
        me = (area / 12.0) * (np.ones((3, 3)) + np.eye(3))
        d = dof[tri_idx[t]]
        for i in range(3):
            for j in range(3):
                if d[i] >= 0 and d[j] >= 0:
                    A[d[i], d[j]] += ke[i, j]
                    B[d[i], d[j]] += me[i, j]
    w, V = sla.eigh(A, B)
    eps, mu = EPS0 * eps_r, MU0 * mu_r
    k = omega * np.sqrt(complex(mu * eps))
    out = []
    # reference: `kc2 < 1e-12` (port_eigensolve.cpp:201).  The TE null mode lands at ~1e-16*lambda_max from a dense
    # eigen-solver, above or below that absolute constant depending on the solver; both sides of the parity test use
    # the scale-aware cut max(1e-12, 1e-9*lambda_max) so the null mode is always dropped.
    null_cut = max(1e-12, 1e-9 * float(np.max(np.abs(w[np.isfinite(w)]))))
    for i in range(nd):
        kc2 = float(w[i])
        if not np.isfinite(kc2) or kc2 < null_cut:
            continue
        kc = np.sqrt(kc2)
        m = PortMode()
        m.pol = 1 if tm else 0
        m.fc, m.kc, m.omega, m.eps, m.mu = kc * C0 / (2 * np.pi), kc, omega, eps, mu
        beta = np.sqrt(complex(k * k - kc2))
        m.beta = beta
        m.Z0 = 0.0 if beta == 0 else ((omega * mu / beta) if not tm else (beta / (omega * eps)))
        f = np.zeros(nn, dtype=np.complex128)
        f[~pec] = V[:, i]
        if f[np.argmax(np.abs(f))].real < 0:
            f = -f
        power = 0.0 + 0.0j
        for t in range(tri_idx.shape[0]):
            G, area = shapes[t]
            g = G.T @ f[tri_idx[t]]
            if not tm:
                fe, fh = 1j * omega * mu / kc2, beta / kc2
                Ex, Ey, Hx, Hy = -fe * g[1], fe * g[0], fh * g[0], fh * g[1]
            else:
                fe, fh = -beta / kc2, 1.0 / (1j * omega * mu)
                Ex, Ey, Hx, Hy = fe * g[0], fe * g[1], -fh * g[1], fh * g[0]
            power += 0.5 * area * (Ex * np.conj(Hy) - Ey * np.conj(Hx))
        pr = power.real if power.real > 0 else abs(power)
        if pr <= 0:
            continue
        out.append((m, f / np.sqrt(pr)))
        if len(out) >= num_modes:
            break
    out.sort(key=lambda mf: mf[0].fc)
    return out[:num_modes]


def populate_te10_field(surf: PortSurface, a: float, b: float, mode: PortMode) -> np.ndarray:  # wave_port.cpp:214-262
    x = surf.xy[:, 0] - surf.xy[:, 0].min()
    A_sq = 4.0 * mode.kc ** 4 * a / (mode.omega * np.real(mode.mu) * np.real(mode.beta) * np.pi ** 2 * b)
    return (np.cos(np.pi * x / a) * np.sqrt(A_sq)).astype(np.complex128)


def build_wave_port(mesh: Mesh, surf: PortSurface, mode: PortMode, fld: np.ndarray) -> WavePort:  # wave_port.cpp:114-212
    if fld.size != surf.node_ids.size:
        raise RuntimeError("Port mode field size does not match surface mesh nodes")
    accum: Dict[int, complex] = {}
    counts: Dict[int, int] = {}
    normal = 0.0
    tri_idx = surf.idx(surf.tri_conn)
    for t in range(tri_idx.shape[0]):
        G, _ = _tri_shape(surf.xy[tri_idx[t]])
        g = G.T @ fld[tri_idx[t]]
        if mode.kc == 0.0:
            raise RuntimeError("Port mode has zero cutoff wavenumber")
        if mode.pol == 0:
            f = 1j * mode.omega * mode.mu / (mode.kc ** 2)
            Ex, Ey = f * g[1], -f * g[0]
        else:
            f = -mode.beta / (mode.kc ** 2)
            Ex, Ey = f * g[0], f * g[1]
        vt = int(surf.volume_tri_indices[t])
        P = mesh.xyz[mesh.node_idx_of(mesh.tri_conn[vt])]
        normal += np.cross(P[1] - P[0], P[2] - P[0])[2]
        for e in range(3):
            ei = int(mesh.tri_edges[vt, e])
            pa, pb = mesh.xyz[mesh.node_idx_of(mesh.edges[ei])]
            ev = pb - pa
            accum[ei] = accum.get(ei, 0.0) + Ex * ev[0] + Ey * ev[1]
            counts[ei] = counts.get(ei, 0) + 1
    sign = 1.0 if normal >= 0 else -1.0
    edges = sorted(accum)
    port = WavePort()
    port.surface_tag = surf.phys
    port.mode = mode
    port.edges = edges
    port.weights = np.array([sign * accum[e] / counts[e] for e in edges], dtype=np.complex128)
    return port


def build_wave_port_from_eigenvector(mesh: Mesh, surf: PortSurface, ev: np.ndarray, mode: PortMode, pec: Set[int]) -> WavePort:
    """src/ports/wave_port.cpp:264-309: weights = j * eigenvector on the port's free edges, ||w||^2 = sqrt(Re Z0)."""
    edges = sorted({int(e) for t in surf.volume_tri_indices for e in mesh.tri_edges[int(t)]})
    w = np.array([0.0 if e in pec else 1j * ev[e] for e in edges], dtype=np.complex128)
    norm_sq = float(sum(ev[e] ** 2 for e in edges if e not in pec))
    if norm_sq > 1e-15 and np.real(mode.Z0) > 1e-15:
        w = w * np.sqrt(np.sqrt(np.real(mode.Z0)) / norm_sq)
    port = WavePort()
    port.surface_tag, port.mode, port.edges, port.weights = surf.phys, mode, edges, w
    return port


def straight_waveguide_sparams(a: float, length: float, freq: float):  # port_eigensolve.cpp:67-88 -> (s11, s21, s12, s22)
    k, kc = 2 * np.pi * freq / C0, np.pi / a
    if k <= kc:
        return 1.0 + 0j, 0j, 0j, 1.0 + 0j
    ph = np.exp(-1j * np.sqrt(k * k - kc * kc) * length)
    return 0j, ph, ph, 0j


def _g12(x: float) -> str:
    """C++ ostream << double with setprecision(12) (default float field = %g)."""
    return "%.12g" % x


def touchstone_nport_text(freq, S_list, fmt: str = "RI", z0: float = 50.0) -> str:  # src/io/touchstone.cpp:65-139
    def val(v):
        if fmt == "RI":
            return _g12(v.real) + " " + _g12(v.imag)
        ang = np.angle(v) * 180.0 / np.pi
        if fmt == "MA":
            return _g12(abs(v)) + " " + _g12(ang)
        return _g12(20.0 * np.log10(max(abs(v), 1e-20))) + " " + _g12(ang)

    n = S_list[0].shape[0]
    out = ["! Touchstone file generated by EdgeFEM\n", "! Number of ports: %d\n" % n, "# Hz S %s R %s\n" % (fmt, _g12(z0))]
    for f, S in zip(freq, S_list):
        line = _g12(f)
        cnt = 0
        for i in range(n):
            for j in range(n):
                if n > 2 and cnt > 0 and cnt % 4 == 0:
                    line += "\n"
                line += " " + val(complex(S[i, j]))
                cnt += 1
        out.append(line + "\n")
    return "".join(out)


def touchstone_legacy_text(freq, sp) -> str:  # src/io/touchstone.cpp:50-62; sp = list of (s11, s21, s12, s22)
    out = ["# Hz S RI R 50\n"]
    for f, s in zip(freq, sp):
        out.append(" ".join([_g12(f)] + [_g12(x) for v in s for x in (complex(v).real, complex(v).imag)]) + "\n")
    return "".join(out)


# =====================================================================================================
# field post-processing (test infrastructure)
# =====================================================================================================
_EDGE_PAIRS = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


def compute_barycentric(X: np.ndarray, p: np.ndarray) -> np.ndarray:  # src/edge_basis.cpp:132-151
    T = np.column_stack([X[0] - X[3], X[1] - X[3], X[2] - X[3]])
    l = np.linalg.inv(T) @ (np.asarray(p, dtype=float) - X[3])
    return np.array([l[0], l[1], l[2], 1.0 - l[0] - l[1] - l[2]])


def whitney_edge_curls(X: np.ndarray) -> np.ndarray:  # src/edge_basis.cpp:33-46 -> [6,3]
    g, _ = gradients_and_volume(X)
    return np.array([2.0 * np.cross(g[a], g[b]) for a, b in _EDGE_PAIRS])


def evaluate_edge_field(X: np.ndarray, orient, dofs, p) -> np.ndarray:  # src/edge_basis.cpp:161-190
    lam = compute_barycentric(X, p)
    g, _ = gradients_and_volume(X)
    E = np.zeros(3, dtype=np.complex128)
    for e, (a, b) in enumerate(_EDGE_PAIRS):
        E += dofs[e] * float(orient[e]) * (lam[a] * g[b] - lam[b] * g[a])
    return E


def extract_huygens_surface(mesh: Mesh, x: np.ndarray, surface_tag: int, omega: float, mu_r: complex = 1.0):
    """src/post/huygens_surface.cpp:29-149 -> dict(r, n, E_tan, H_tan, area).  Eigen's `a.dot(b)` conjugates a, so the
    'tangential' projection subtracts conj(E.n) n (restated as written)."""
    jwmu = 1j * omega * (4.0e-7 * np.pi) * mu_r
    faces = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))
    tri_to_tet: Dict[Tuple[int, int, int], int] = {}
    for t in range(mesh.tet_conn.shape[0]):
        c = mesh.tet_conn[t]
        for f in faces:
            tri_to_tet[tuple(sorted((int(c[f[0]]), int(c[f[1]]), int(c[f[2]]))))] = t  # later tets overwrite
    out = dict(r=[], n=[], E_tan=[], H_tan=[], area=[])
    for i in range(mesh.tri_conn.shape[0]):
        if mesh.tri_phys[i] != surface_tag:
            continue
        v = mesh.xyz[mesh.node_idx_of(mesh.tri_conn[i])]
        cen = (v[0] + v[1] + v[2]) / 3.0
        cr = np.cross(v[1] - v[0], v[2] - v[0])
        area = 0.5 * np.linalg.norm(cr)
        if area < 1e-30:
            continue
        normal = cr / np.linalg.norm(cr)
        t = tri_to_tet.get(tuple(sorted(int(q) for q in mesh.tri_conn[i])))
        if t is None:
            continue
        X = mesh.xyz[mesh.node_idx_of(mesh.tet_conn[t])]
        if normal @ (cen - X.mean(axis=0)) < 0:
            normal = -normal
        dofs = x[mesh.tet_edges[t]]
        E = evaluate_edge_field(X, mesh.tet_orient[t], dofs, cen)
        curl = (dofs * mesh.tet_orient[t].astype(float)) @ whitney_edge_curls(X)
        H = curl / jwmu
        E_tan = E - np.conj(E @ normal) * normal
        H_tan = H - np.conj(H @ normal) * normal
        for k, val in zip(("r", "n", "E_tan", "H_tan", "area"), (cen, normal, E_tan, H_tan, area)):
            out[k].append(val)
    if not out["r"]:
        raise RuntimeError("extract_huygens_surface: no triangles found with surface_tag = %d" % surface_tag)
    return {k: np.array(v) for k, v in out.items()}


Z0_FREE = 376.730313668  # src/post/ntf.cpp:11


def stratton_chu(r, n, E, H, area, theta, phi, k0):
    """E_theta, E_phi at the direction list (theta[i], phi[i]); src/post/ntf.cpp:86-203 (2-D and 3-D share the kernel)."""
    r, n, E, H, area = (np.asarray(a) for a in (r, n, E, H, area))
    et, ep = np.zeros(len(theta), dtype=np.complex128), np.zeros(len(theta), dtype=np.complex128)
    J = np.cross(n, H)
    M = -np.cross(n, E)
    for i, (th, ph) in enumerate(zip(theta, phi)):
        rhat = np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        th_hat = np.array([np.cos(th) * np.cos(ph), np.cos(th) * np.sin(ph), -np.sin(th)])
        ph_hat = np.array([-np.sin(ph), np.cos(ph), 0.0])
        phase = np.exp(-1j * k0 * (r @ rhat))
        term = Z0_FREE * np.cross(np.cross(rhat, J), rhat) - np.cross(rhat, M)
        Efar = ((1j * k0 / (4 * np.pi)) * term * (phase * area)[:, None]).sum(axis=0)
        et[i], ep[i] = Efar @ th_hat, Efar @ ph_hat
    return et, ep


def compute_directivity(theta, phi, E_theta, E_phi) -> float:  # src/post/ntf.cpp:209-262 (theta rows, phi columns)
    pwr = np.abs(E_theta) ** 2 + np.abs(E_phi) ** 2
    U_max = pwr.max() / (2 * Z0_FREE)
    if U_max < np.finfo(float).eps:
        return 0.0
    Nth, Nph = pwr.shape
    if Nth < 2 or Nph < 2:
        return 1.0
    dth, dph = (theta[-1] - theta[0]) / (Nth - 1), (phi[-1] - phi[0]) / (Nph - 1)
    wt, wp = np.ones(Nth), np.ones(Nph)
    wt[[0, -1]] = 0.5
    wp[[0, -1]] = 0.5
    P = ((pwr / (2 * Z0_FREE)) * (np.sin(theta) * wt)[:, None] * wp[None, :]).sum() * dth * dph
    return 0.0 if P < np.finfo(float).eps else float(4 * np.pi * U_max / P)
