"""CPU check of the cluster split (edgefem_b200/csrc/cluster_plan.hpp) that k_cocg_cluster runs on: the plan is built by
the library's host code (efb_debug_cluster_plan_*, no GPU needed) and the kernel's algorithm -- thread-per-row ELL
SpMV out of a window, halo pulls from the owners, nodal partial exchange, G^T r by recurrence, partials summed in rank
order -- is replayed here in numpy, phase by phase over the CTAs, against a direct solve."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import edgefem_oracle as orc
from edgefem_b200 import cabi

INFO = dict(N_OWN=0, LO=1, WLO=2, WN=3, N_MY=4, N_HALO=5, N_BLK=6, N_SLOTS=7, OFF_ROW=8, OFF_SLOT=9, OFF_BLK=10, OFF_HALO=11,
            OFF_NODE=12, OFF_N2E=13, OFF_NSRC=14, N_PUSH=15, OFF_PUSH=16)
STRIDE = 20


class Emu:
    """One job of k_cocg_cluster<NR=1> replayed on the host."""

    def __init__(self, plan, A_csr, dinv, linv, aux):
        self.p, self.A, self.aux = plan, A_csr, aux
        d = plan["dims"]
        self.C = int(d[0])
        self.info = plan["cta_info"].reshape(self.C, STRIDE)
        self.vals = A_csr.data
        self.dinv, self.linv = dinv, linv

    def I(self, c, k):
        return int(self.info[c, INFO[k]])

    def rows(self, c, name):
        o, n = self.I(c, "OFF_ROW"), self.I(c, "N_OWN")
        return self.p[name][o:o + n]

    def load(self):
        self.mat = []
        for c in range(self.C):
            o, n = self.I(c, "OFF_SLOT"), self.I(c, "N_SLOTS")
            src = self.p["slot_src"][o:o + n]
            v = np.where(src >= 0, self.vals[np.maximum(src, 0)], 0.0).astype(complex)
            # blocks of all-real rows keep the real part only (8-byte values in shared memory)
            ob = self.I(c, "OFF_BLK")
            for b in range(self.I(c, "N_BLK")):
                o0, o1 = int(self.p["blk_off"][ob + b]), int(self.p["blk_off"][ob + b + 1])
                v0, v1 = int(self.p["blk_voff"][ob + b]), int(self.p["blk_voff"][ob + b + 1])
                assert v1 - v0 in (o1 - o0, 2 * (o1 - o0))
                if v1 - v0 != 2 * (o1 - o0):
                    v[o0:o1] = v[o0:o1].real
                else:
                    assert v0 % 2 == 0  # complex128 values stay 16-byte aligned
            self.mat.append((v, self.p["slot_col"][o:o + n]))

    def spmv(self, c, p_w):
        n_own, n_blk, ob = self.I(c, "N_OWN"), self.I(c, "N_BLK"), self.I(c, "OFF_BLK")
        v, col = self.mat[c]
        out = np.zeros(n_own, dtype=complex)
        for b in range(n_blk):
            base, end = int(self.p["blk_off"][ob + b]), int(self.p["blk_off"][ob + b + 1])
            w = (end - base) // 32
            blk_v = v[base:end].reshape(w, 32)
            blk_c = col[base:end].reshape(w, 32)
            acc = (blk_v * p_w[blk_c]).sum(axis=0)
            n = min(32, n_own - 32 * b)
            out[32 * b:32 * b + n] = acc[:n]
        return out

    def pull_halo(self, c, p_w, z_own, beta=None):
        o, n = self.I(c, "OFF_HALO"), self.I(c, "N_HALO")
        for h in range(n):
            hw, src = int(self.p["halo_ws"][o + h]), int(self.p["halo_src"][o + h])
            z = z_own[src >> 16][src & 0xffff]
            p_w[hw] = z if beta is None else z + beta * p_w[hw]

    def nodal_partial(self, c, q_own):
        n_my, on, oe = self.I(c, "N_MY"), self.I(c, "OFF_NODE") + c, self.I(c, "OFF_N2E")
        wp = np.zeros(n_my, dtype=complex)
        for j in range(n_my):
            for k in range(int(self.p["n2e_ptr"][on + j]), int(self.p["n2e_ptr"][on + j + 1])):
                it = int(self.p["n2e_item"][oe + k])
                wp[j] += q_own[it >> 1] if it & 1 else -q_own[it >> 1]
        return wp

    def nodal_combine(self, c, wps):
        n_my, on, os_ = self.I(c, "N_MY"), self.I(c, "OFF_NODE") + c, self.I(c, "OFF_NSRC")
        out = np.zeros(n_my, dtype=complex)
        for j in range(n_my):
            for k in range(int(self.p["nsrc_ptr"][on + j]), int(self.p["nsrc_ptr"][on + j + 1])):
                it = int(self.p["nsrc_item"][os_ + k])
                out[j] += wps[it >> 16][it & 0xffff]
        return out

    def solve(self, b, tol=1e-10, max_it=5000):
        C = self.C
        self.load()
        edge = [self.rows(c, "row_edge") for c in range(C)]
        ws = [self.rows(c, "row_ws") for c in range(C)]
        n0 = [self.rows(c, "row_n0") for c in range(C)]
        n1 = [self.rows(c, "row_n1") for c in range(C)]
        di = [self.dinv[edge[c]] for c in range(C)]
        li = []
        for c in range(C):
            o, n = self.I(c, "OFF_NODE"), self.I(c, "N_MY")
            li.append(self.linv[self.p["node_id"][o:o + n]] if self.aux else np.zeros(0, dtype=complex))
        x = [np.zeros(len(edge[c]), dtype=complex) for c in range(C)]
        r = [b[edge[c]].astype(complex) for c in range(C)]
        p_w = [np.zeros(self.I(c, "WN"), dtype=complex) for c in range(C)]
        bb = sum(float(np.vdot(v, v).real) for v in r)
        # g = G^T r, w = linv g
        wps = [self.nodal_partial(c, r[c]) for c in range(C)]
        g = [self.nodal_combine(c, wps) for c in range(C)]

        def precond():
            z = []
            for c in range(C):
                zc = di[c] * r[c]
                if self.aux:
                    w = li[c] * g[c]
                    zc = zc + w[n1[c]] - w[n0[c]]
                z.append(zc)
            return z

        z = precond()
        rho = sum((r[c] * z[c]).sum() for c in range(C))
        pr = [z[c].copy() for c in range(C)]
        for c in range(C):
            p_w[c][ws[c]] = pr[c]
        for c in range(C):
            self.pull_halo(c, p_w[c], z)
        it = 0
        while it < max_it:
            q = [self.spmv(c, p_w[c]) for c in range(C)]
            pq = sum((pr[c] * q[c]).sum() for c in range(C))
            wps = [self.nodal_partial(c, q[c]) for c in range(C)]
            alpha = rho / pq
            it += 1
            for c in range(C):
                x[c] += alpha * pr[c]
                r[c] -= alpha * q[c]
                if self.aux:
                    g[c] = g[c] - alpha * self.nodal_combine(c, wps)
            z = precond()
            rho_new = sum((r[c] * z[c]).sum() for c in range(C))
            rr = sum(float(np.vdot(v, v).real) for v in r)
            if rr <= tol * tol * bb:
                break
            beta = rho_new / rho
            rho = rho_new
            for c in range(C):
                pr[c] = z[c] + beta * pr[c]
                p_w[c][ws[c]] = pr[c]
            for c in range(C):
                self.pull_halo(c, p_w[c], z, beta)
        xf = np.zeros(self.A.shape[0], dtype=complex)
        for c in range(C):
            xf[edge[c]] = x[c]
        return xf, it


def wr90_system(wr90):
    mesh, pec = wr90
    p = orc.MaxwellParams(omega=2 * math.pi * 10e9)
    ports = orc.wr90_ports(mesh, pec, 10e9)
    A, port_mass, port_vecs, betas = orc.eigenmode_system(mesh, p, pec, ports)
    bs = [2.0 * 1j * betas[a] * (port_mass[a] @ port_vecs[a]) for a in range(len(ports))]  # src/assemble_maxwell.cpp:749-752
    A = sp.csr_matrix(A)
    A.sort_indices()
    return mesh, pec, A, bs


def precond_arrays(mesh, A, pm):
    m = A.shape[0]
    en = mesh.node_idx_of(mesh.edges).astype(np.int32)
    nn = mesh.xyz.shape[0]
    free = np.nonzero(~pm)[0]
    G = sp.csr_matrix((np.concatenate([-np.ones(free.size), np.ones(free.size)]), (np.concatenate([free, free]), np.concatenate([en[free, 0], en[free, 1]]))),
                      shape=(m, nn))
    nd = np.zeros(nn, bool)
    nd[en[pm].reshape(-1)] = True
    L = (G.T @ A @ G).diagonal()
    linv = np.where((~nd) & (L != 0), 1.0 / np.where(L == 0, 1, L), 0)
    return en, 1.0 / A.diagonal(), linv


def complex_rows(A):
    """Flags of the rows with a value that is not real (what k_row_complex computes on the device)."""
    flag = np.zeros(A.shape[0], np.uint8)
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    flag[rows[A.data.imag != 0]] = 1
    return flag


@pytest.mark.parametrize("C,mixed", [(1, False), (2, False), (4, False), (8, False), (3, True), (6, True), (8, True)])
@pytest.mark.parametrize("aux", [True, False])
def test_cluster_plan_emulation_matches_direct_solve(wr90, C, mixed, aux):
    if not aux and (C, mixed) not in ((1, False), (8, False), (6, True)):
        pytest.skip("Jacobi needs ~1800 iterations: three shapes are enough")
    mesh, pec, A, bs = wr90_system(wr90)
    pm = orc.pec_mask(mesh, pec)
    en, dinv, linv = precond_arrays(mesh, A, pm)
    rc = complex_rows(A) if mixed else None
    if mixed:  # the lossless guide is real except for the rows of the two port faces
        assert 0 < int(rc[~pm].sum()) < 0.1 * int((~pm).sum())
    plan = cabi.cluster_plan_arrays(A.indptr, A.indices, pm.astype(np.uint8), mesh.xyz.shape[0] if aux else 0, en if aux else None, C, rc)
    d = plan["dims"]
    assert d[0] == C and d[1] == int((~pm).sum())
    info = plan["cta_info"].reshape(C, STRIDE)
    assert info[:, INFO["N_OWN"]].sum() == d[1]
    # every free edge is owned exactly once; windows contain their own rows
    assert sorted(plan["row_edge"].tolist()) == np.nonzero(~pm)[0].tolist()
    # the push lists (what the owners send before the barrier) are exactly the halo lists (what the readers need) transposed
    want = set()
    for c in range(C):
        o, n = int(info[c, INFO["OFF_HALO"]]), int(info[c, INFO["N_HALO"]])
        for h in range(n):
            src = int(plan["halo_src"][o + h])
            want.add((src >> 16, src & 0xffff, c, h))
    got = set()
    for c in range(C):
        o, n = int(info[c, INFO["OFF_PUSH"]]), int(info[c, INFO["N_PUSH"]])
        for i in range(n):
            dst = int(plan["push_dst"][o + i])
            got.add((c, int(plan["push_row"][o + i]), dst >> 16, dst & 0xffff))
    assert got == want and len(got) == int(info[:, INFO["N_PUSH"]].sum())
    b = np.asarray(bs[0]).astype(complex)
    xe, it = Emu(plan, A, dinv, linv, aux).solve(b, tol=1e-10, max_it=6000)
    xd = spla.splu(sp.csc_matrix(A)).solve(b)
    free = ~pm
    assert it < (600 if aux else 4000)
    err = np.linalg.norm(xe[free] - xd[free]) / np.linalg.norm(xd[free])
    assert err < 1e-7, (C, aux, it, err)


def test_cluster_plan_band_and_footprint(wr90):
    """RCM keeps the WR-90 matrix banded: the window of a CTA stays a small multiple of its own rows, and the 8-CTA split
    of one right-hand side fits the 227 KB of shared memory of an SM."""
    mesh, pec, A, bs = wr90_system(wr90)
    pm = orc.pec_mask(mesh, pec)
    en, dinv, linv = precond_arrays(mesh, A, pm)
    plan = cabi.cluster_plan_arrays(A.indptr, A.indices, pm.astype(np.uint8), mesh.xyz.shape[0], en, 8)
    info = plan["cta_info"].reshape(8, STRIDE)
    own, wn = info[:, INFO["N_OWN"]], info[:, INFO["WN"]]
    assert own.max() <= 640 and (wn <= 2.2 * own).all()
    assert plan["dims"][9] <= 227 * 1024  # NR = 1
    # padding of the ELL blocks stays small (rows sorted by length inside a CTA)
    nnz_free = int(np.count_nonzero((~pm)[np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))] & (~pm)[A.indices]))
    assert info[:, INFO["N_SLOTS"]].sum() <= 1.15 * nnz_free
    assert plan["dims"][11] == 2 * info[:, INFO["N_SLOTS"]].max()  # every block complex: 2 units per slot
    # with the real rows stored as doubles the split over SIX CTAs fits too (24 clusters resident instead of 15) ...
    plan6 = cabi.cluster_plan_arrays(A.indptr, A.indices, pm.astype(np.uint8), mesh.xyz.shape[0], en, 6, complex_rows(A))
    assert plan6["dims"][9] <= 227 * 1024 and plan6["dims"][4] <= 2 * 640  # two rows per thread
    # ... which the all-complex storage does not
    plan6c = cabi.cluster_plan_arrays(A.indptr, A.indices, pm.astype(np.uint8), mesh.xyz.shape[0], en, 6)
    assert plan6c["dims"][9] > 227 * 1024


def test_cluster_plan_tiny_and_no_dirichlet():
    """Ragged corner cases: fewer rows than CTAs, no Dirichlet flags, no gradient."""
    n = 5
    A = sp.diags([np.full(n - 1, -1.0), np.full(n, 4.0), np.full(n - 1, -1.0)], [-1, 0, 1], format="csr").astype(complex)
    plan = cabi.cluster_plan_arrays(A.indptr, A.indices, None, 0, None, 8)
    assert plan["dims"][1] == n
    b = np.arange(1, n + 1).astype(complex)
    x, it = Emu(plan, A, 1.0 / A.diagonal(), np.zeros(0), False).solve(b, tol=1e-12, max_it=50)
    assert np.allclose(A @ x, b, atol=1e-10)


@pytest.mark.parametrize("C", [2, 5, 7])
def test_cluster_plan_structure_on_a_layered_unit_cell(C):
    """Structural invariants of the plan on another mesh (layered unit cell, PEC ground + patch, lossy substrate so that most
    rows are complex): every free edge owned once, windows cover the rows' columns, every own edge appears at both of its
    end nodes, halo and push lists are each other's transpose, value offsets follow the block kinds."""
    from edgefem_b200 import meshgen

    xyz, tets, tp, tris, trp = meshgen.unit_cell(px=5e-3, py=5e-3, h_sub=0.5e-3, h_air=6e-3, nx=5, ny=5, nz_sub=2, nz_air=4, patch=(3e-3, 3e-3))
    mesh = orc.mesh_from_arrays(xyz, tets, tp, tris, trp)
    pec = orc.build_edge_pec(mesh, 1) | orc.build_edge_pec(mesh, 2)
    p = orc.MaxwellParams(omega=2 * math.pi * 10e9, eps_r_regions={100: complex(3.5, -0.1), 101: complex(1.0, 0.0)})
    A = sp.csr_matrix(orc.assemble_maxwell(mesh, p, pec, [])[0])
    A.sort_indices()
    pm = orc.pec_mask(mesh, pec)
    en = mesh.node_idx_of(mesh.edges).astype(np.int32)
    rc = complex_rows(A)
    assert 0 < int(rc[~pm].sum()) < int((~pm).sum())  # substrate rows complex, air rows real
    plan = cabi.cluster_plan_arrays(A.indptr, A.indices, pm.astype(np.uint8), mesh.xyz.shape[0], en, C, rc)
    info = plan["cta_info"].reshape(C, STRIDE)
    free = np.nonzero(~pm)[0]
    assert sorted(plan["row_edge"].tolist()) == free.tolist()
    c_orig = plan["c_orig"]
    pos_of = {int(e): i for i, e in enumerate(c_orig)}
    seen_push, seen_halo = set(), set()
    for c in range(C):
        I = {k: int(info[c, v]) for k, v in INFO.items()}
        rows = plan["row_edge"][I["OFF_ROW"]:I["OFF_ROW"] + I["N_OWN"]]
        ws = plan["row_ws"][I["OFF_ROW"]:I["OFF_ROW"] + I["N_OWN"]]
        # the window slot of an own row is its position minus the window start
        assert all(pos_of[int(e)] - I["WLO"] == int(w) for e, w in zip(rows, ws))
        # blocks: slots of row t sit at lane t % 32 of block t // 32; their columns are window slots of the row's free columns
        for t, e in enumerate(rows):
            b, l = divmod(t, 32)
            o0, o1 = int(plan["blk_off"][I["OFF_BLK"] + b]), int(plan["blk_off"][I["OFF_BLK"] + b + 1])
            src = plan["slot_src"][I["OFF_SLOT"] + o0 + l:I["OFF_SLOT"] + o1:32]
            col = plan["slot_col"][I["OFF_SLOT"] + o0 + l:I["OFF_SLOT"] + o1:32]
            used = src >= 0
            want_cols = [j for j in A.indices[A.indptr[e]:A.indptr[e + 1]] if not pm[j]]
            assert sorted(int(c_orig[I["WLO"] + int(w)]) for w in col[used]) == sorted(int(j) for j in want_cols)
            assert sorted(int(A.indices[k]) for k in src[used]) == sorted(int(j) for j in want_cols)
            v0, v1 = int(plan["blk_voff"][I["OFF_BLK"] + b]), int(plan["blk_voff"][I["OFF_BLK"] + b + 1])
            if rc[e]:
                assert v1 - v0 == 2 * (o1 - o0)  # a complex row lives in a complex block
        # nodal lists: every own row appears once as tail (bit 0 clear) and once as head (bit 0 set)
        n_my = I["N_MY"]
        ptr = plan["n2e_ptr"][I["OFF_NODE"] + c:I["OFF_NODE"] + c + n_my + 1]
        items = plan["n2e_item"][I["OFF_N2E"]:I["OFF_N2E"] + int(ptr[-1])]
        assert sorted(items.tolist()) == sorted([2 * t for t in range(I["N_OWN"])] + [2 * t + 1 for t in range(I["N_OWN"])])
        for h in range(I["N_HALO"]):
            src = int(plan["halo_src"][I["OFF_HALO"] + h])
            seen_halo.add((src >> 16, src & 0xffff, c, h))
        for i in range(I["N_PUSH"]):
            dst = int(plan["push_dst"][I["OFF_PUSH"] + i])
            seen_push.add((c, int(plan["push_row"][I["OFF_PUSH"] + i]), dst >> 16, dst & 0xffff))
    assert seen_push == seen_halo
