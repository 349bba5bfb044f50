"""ctypes declarations for the C-ABI in include/edgefem_b200.h (libedgefem_b200.so).

Product code: a thin, typed view of the shared library -- no arithmetic happens here and
nothing from ``oracle/`` is imported.  The C++ host layer (``pyedgefem``) is the intended
user-facing API; this module exists for the C-ABI parity tests, ``bench.py`` kernel timing
and ``__graft_entry__.smoke()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libedgefem_b200.so")

EFB_OK = 0
METHOD_AUTO, METHOD_BICGSTAB, METHOD_COCG, METHOD_DIRECT = 0, 1, 2, 3
PRECOND_JACOBI, PRECOND_AUX, PRECOND_NONE = 0, 1, 2
MODEL_NONE, MODEL_DEBYE, MODEL_LORENTZ, MODEL_DRUDE, MODEL_DRUDE_LORENTZ = 0, 1, 2, 3, 4
PML_NONE, PML_UNIFORM, PML_TENSOR = 0, 1, 2

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)


class MeshDesc(C.Structure):
    _fields_ = [
        ("n_node", C.c_int32), ("xyz", f64p), ("n_tet", C.c_int32), ("tet_nodes", i32p), ("tet_edges", i32p),
        ("tet_orient", i8p), ("tet_phys", i32p), ("n_edge", C.c_int32), ("edge_nodes", i32p),
    ]


class Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_poles", C.c_int32), ("pole_begin", C.c_int32), ("_pad", C.c_int32),
                ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double)]


class Pole(C.Structure):
    _fields_ = [("delta_eps", C.c_double), ("omega0", C.c_double), ("gamma", C.c_double)]


class Pml(C.Structure):
    _fields_ = [("kind", C.c_int32), ("enforce_heuristics", C.c_int32), ("sigma", C.c_double * 3),
                ("thickness", C.c_double * 3), ("grading_order", C.c_double)]


class Materials(C.Structure):
    _fields_ = [("n_slots", C.c_int32), ("eps_static", f64p), ("mu_static", f64p), ("eps_models", C.POINTER(Model)),
                ("mu_models", C.POINTER(Model)), ("n_poles", C.c_int32), ("poles", C.POINTER(Pole)), ("pml", C.POINTER(Pml))]


class SolveOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("precond", C.c_int32), ("tolerance", C.c_double), ("max_iterations", C.c_int32),
                ("check_every", C.c_int32), ("symmetric_hint", C.c_int32), ("zero_initial_guess", C.c_int32),
                ("max_restarts", C.c_int32), ("_pad", C.c_int32)]


class SolveResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32), ("method", C.c_int32), ("precond", C.c_int32),
                ("residual", C.c_double)]


# name -> (restype, argtypes); this table is also what tests/test_cabi_symbols.py checks against the header
SIGNATURES = {
    "efb_abi_version": (C.c_int, []),
    "efb_device_count": (C.c_int, []),
    "efb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "efb_ctx_destroy": (None, [C.c_void_p]),
    "efb_last_error": (C.c_char_p, [C.c_void_p]),
    "efb_ctx_sync": (C.c_int, [C.c_void_p]),
    "efb_l2_flush": (C.c_int, [C.c_void_p]),
    "efb_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "efb_launch_count": (C.c_int64, [C.c_void_p]),
    "efb_timer_start": (C.c_int, [C.c_void_p]),
    "efb_timer_stop": (C.c_int, [C.c_void_p, f64p]),
    "efb_mesh_create": (C.c_int, [C.c_void_p, C.POINTER(MeshDesc), C.POINTER(C.c_void_p)]),
    "efb_mesh_destroy": (None, [C.c_void_p]),
    "efb_mesh_num_slots": (C.c_int, [C.c_void_p]),
    "efb_mesh_get_slot_tags": (C.c_int, [C.c_void_p, i32p]),
    "efb_system_create": (C.c_int, [C.c_void_p, C.c_int64, i32p, i32p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "efb_system_create_csr": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, i32p, i32p, f64p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "efb_system_destroy": (None, [C.c_void_p]),
    "efb_system_dims": (C.c_int, [C.c_void_p, i32p, i64p, i32p, i32p]),
    "efb_system_get_pattern": (C.c_int, [C.c_void_p, i32p, i32p]),
    "efb_system_get_values": (C.c_int, [C.c_void_p, C.c_int32, f64p]),
    "efb_system_set_values": (C.c_int, [C.c_void_p, C.c_int32, f64p]),
    "efb_system_set_dirichlet": (C.c_int, [C.c_void_p, u8p]),
    "efb_system_set_gradient": (C.c_int, [C.c_void_p, C.c_int32, i32p]),
    "efb_assemble_volume": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, f64p, C.POINTER(Materials), C.c_int32]),
    "efb_combine_km": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, f64p, C.c_int32, C.c_int32]),
    "efb_add_diag": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, i32p, f64p]),
    "efb_port_create": (C.c_int, [C.c_void_p, C.c_int32, i32p, f64p, C.c_int64, i32p, i32p, f64p, C.POINTER(C.c_void_p)]),
    "efb_port_destroy": (None, [C.c_void_p]),
    "efb_port_normalize_mass": (C.c_int, [C.c_void_p, f64p]),
    "efb_port_add_block": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, f64p]),
    "efb_port_add_mass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, f64p]),
    "efb_port_rhs_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, f64p]),
    "efb_port_rhs_mass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, f64p]),
    "efb_port_project_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, f64p]),
    "efb_port_project_mass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, f64p]),
    "efb_port_rhs_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, i32p, f64p, C.c_int32]),
    "efb_port_project_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, i32p, f64p, C.c_int32]),
    "efb_periodic_extra": (C.c_int, [C.c_void_p, C.c_int64, i32p, i32p, C.c_int32, i32p, i32p, i64p, i32p, i32p]),
    "efb_apply_periodic": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, f64p]),
    "efb_rhs_zero": (C.c_int, [C.c_void_p, C.c_int32]),
    "efb_rhs_set": (C.c_int, [C.c_void_p, C.c_int32, f64p]),
    "efb_rhs_get": (C.c_int, [C.c_void_p, C.c_int32, f64p]),
    "efb_x_get": (C.c_int, [C.c_void_p, C.c_int32, f64p]),
    "efb_x_set": (C.c_int, [C.c_void_p, C.c_int32, f64p]),
    "efb_x_recover": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, i32p, i32p, f64p]),
    "efb_solve": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(SolveOpts), C.POINTER(SolveResult)]),
    "efb_solve_direct": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(SolveResult)]),
    "efb_solve_direct_limit": (C.c_int, []),
    "efb_clear_caches": (None, []),
    "efb_system_last_solve_kernel_ms": (C.c_int, [C.c_void_p, f64p]),
    "efb_system_last_solve_shape": (C.c_int, [C.c_void_p, i32p, i32p, i32p]),
    "efb_debug_cluster_plan_build": (C.c_int, [C.c_int32, i32p, i32p, u8p, C.c_int32, i32p, u8p, C.c_int32, C.POINTER(C.c_void_p)]),
    "efb_debug_cluster_plan_free": (None, [C.c_void_p]),
    "efb_debug_cluster_plan_get": (C.c_int64, [C.c_void_p, C.c_char_p, i64p, C.c_int64]),
    "efb_spmv_host": (C.c_int, [C.c_void_p, C.c_int32, f64p, f64p]),
    "efb_bench_kernel": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, f64p]),
    "efb_huygens_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, f64p, C.c_int32, i32p, i32p, i32p, C.c_double, f64p, f64p, f64p, f64p, f64p, f64p]),
    "efb_stratton_chu": (C.c_int, [C.c_void_p, C.c_int32, f64p, f64p, f64p, f64p, f64p, C.c_int32, f64p, f64p, C.c_double, f64p, f64p]),
    "efb_build_edges": (C.c_int, [C.c_void_p, C.c_int64, i64p, C.c_int64, i64p, i32p, i8p, i32p, i8p, i64p, i64p, C.c_int64]),
    "efb_system_create_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "efb_dist_unique_id": (C.c_int, [u8p]),
    "efb_dist_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, u8p]),
    "efb_dist_finalize": (None, [C.c_void_p]),
    "efb_dist_row_range": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "efb_dist_solve": (C.c_int, [C.c_void_p, C.POINTER(SolveOpts), C.POINTER(SolveResult), C.c_int32]),
    "efb_dist_bench": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, f64p]),
}

_lib: Optional[C.CDLL] = None


class EfbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libedgefem_b200.so (fails loudly if it has not been built -- there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EfbError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _p(a: Optional[np.ndarray], typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


_nccl_preloaded = False


def _preload_nccl():
    """Make the NCCL the library dlopen()s the SAME copy torch links against (the wheel's nvidia/nccl/lib/libnccl.so.2):
    loading the system libnccl first would satisfy torch's later `libnccl.so.2` dependency with an older library."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass  # fall back to whatever libnccl.so.2 the loader finds


def dist_unique_id() -> bytes:
    _preload_nccl()
    buf = np.zeros(128, dtype=np.uint8)
    rc = load().efb_dist_unique_id(_p(buf, u8p))
    if rc:
        raise EfbError("efb_dist_unique_id failed (%d): %s" % (rc, load().efb_last_error(None).decode()))
    return buf.tobytes()


def dist_row_range(m: int, rank: int, world: int):
    a, b = C.c_int32(), C.c_int32()
    rc = load().efb_dist_row_range(m, rank, world, C.byref(a), C.byref(b))
    if rc:
        raise EfbError("efb_dist_row_range: bad arguments")
    return a.value, b.value


def cluster_plan_arrays(rowptr, colidx, dir_flags=None, n_node=0, edge_nodes=None, cluster_ctas=8, row_complex=None) -> dict:
    """Host-only diagnostics: the cluster split of a CSR pattern as a dict of int64 arrays (efb_debug_cluster_plan_*).
    row_complex: flags of the rows whose values are not all real (None = every row)."""
    lib = load()
    rp, ci = _i32(rowptr), _i32(colidx)
    d = None if dir_flags is None else np.ascontiguousarray(np.asarray(dir_flags, dtype=np.uint8))
    en = None if edge_nodes is None else _i32(edge_nodes)
    rc_ = None if row_complex is None else np.ascontiguousarray(np.asarray(row_complex, dtype=np.uint8))
    h = C.c_void_p()
    rc = lib.efb_debug_cluster_plan_build(rp.size - 1, _p(rp, i32p), _p(ci, i32p), _p(d, u8p), int(n_node), _p(en, i32p), _p(rc_, u8p),
                                          int(cluster_ctas), C.byref(h))
    if rc != EFB_OK:
        raise EfbError("efb_debug_cluster_plan_build failed (%d): %s" % (rc, (lib.efb_last_error(None) or b"").decode()))
    out = {}
    try:
        for name in ("dims", "c_orig", "cta_info", "row_edge", "row_ws", "row_n0", "row_n1", "blk_off", "blk_voff", "slot_src", "slot_col", "halo_ws",
                     "halo_src", "push_row", "push_dst", "node_id", "n2e_ptr", "n2e_item", "nsrc_ptr", "nsrc_item"):
            n = lib.efb_debug_cluster_plan_get(h, name.encode(), None, 0)
            buf = np.zeros(max(int(n), 1), dtype=np.int64)
            lib.efb_debug_cluster_plan_get(h, name.encode(), _p(buf, i64p), buf.size)
            out[name] = buf[:int(n)]
    finally:
        lib.efb_debug_cluster_plan_free(h)
    return out


class Ctx:
    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.efb_ctx_create(device, C.byref(h))
        if rc != EFB_OK:
            raise EfbError("efb_ctx_create failed (%d): %s" % (rc, self.lib.efb_last_error(None).decode()))
        self.h = h

    def check(self, rc: int, what: str = ""):
        if rc != EFB_OK:
            raise EfbError("%s failed (%d): %s" % (what, rc, self.lib.efb_last_error(self.h).decode()))

    def sync(self):
        self.check(self.lib.efb_ctx_sync(self.h), "efb_ctx_sync")

    def l2_flush(self):
        """Overwrite a 256 MB scratch buffer on the context's stream (evicts L2; measurement helper)."""
        self.check(self.lib.efb_l2_flush(self.h), "efb_l2_flush")

    def last_kernel_ms(self) -> float:
        return float(self.lib.efb_last_kernel_ms(self.h))

    def launch_count(self) -> int:
        return int(self.lib.efb_launch_count(self.h))

    def timer_start(self):
        self.check(self.lib.efb_timer_start(self.h), "efb_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double()
        self.check(self.lib.efb_timer_stop(self.h, C.byref(ms)), "efb_timer_stop")
        return ms.value

    def dist_init(self, rank: int, world: int, unique_id: bytes):
        """Join the NCCL communicator of a row-partitioned solve (`unique_id` from dist_unique_id() on rank 0)."""
        _preload_nccl()
        buf = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert buf.size == 128
        self.check(self.lib.efb_dist_init(self.h, rank, world, _p(buf, u8p)), "efb_dist_init")

    def close(self):
        if self.h:
            self.lib.efb_ctx_destroy(self.h)
            self.h = None


def stratton_chu(ctx: "Ctx", r, n, E, H, area, theta, phi, k0):
    """Far-field E_theta, E_phi at the direction list (theta[i], phi[i]) from Huygens-surface samples (efb_stratton_chu)."""
    r_, n_ = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3), np.ascontiguousarray(n, dtype=np.float64).reshape(-1, 3)
    E_, H_ = _c128(E).reshape(-1, 3), _c128(H).reshape(-1, 3)
    a_ = np.ascontiguousarray(area, dtype=np.float64).reshape(-1)
    th, ph = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1), np.ascontiguousarray(phi, dtype=np.float64).reshape(-1)
    assert th.size == ph.size and r_.shape[0] == n_.shape[0] == E_.shape[0] == H_.shape[0] == a_.size
    et, ep = np.zeros(th.size, dtype=np.complex128), np.zeros(th.size, dtype=np.complex128)
    ctx.check(ctx.lib.efb_stratton_chu(ctx.h, a_.size, _p(r_, f64p), _p(n_, f64p), _p(E_.view(np.float64), f64p), _p(H_.view(np.float64), f64p),
                                       _p(a_, f64p), th.size, _p(th, f64p), _p(ph, f64p), float(k0), _p(et.view(np.float64), f64p),
                                       _p(ep.view(np.float64), f64p)), "efb_stratton_chu")
    return et, ep


def build_edges_device(ctx: "Ctx", tet_conn, tri_conn=None):
    """Global edge numbering on the GPU (efb_build_edges), bit-exact with the reference's build_edges.
    tet_conn [t,4] / tri_conn [k,3] hold node IDS.  Returns (tet_edges [t,6] i32, tet_orient [t,6] i8,
    tri_edges [k,3] i32, tri_orient [k,3] i8, edges [m,2] i64)."""
    tc = np.ascontiguousarray(np.asarray(tet_conn, dtype=np.int64).reshape(-1, 4))
    rc_ = np.ascontiguousarray(np.asarray(tri_conn if tri_conn is not None else np.zeros((0, 3)), dtype=np.int64).reshape(-1, 3))
    nt, nk = tc.shape[0], rc_.shape[0]
    te, to = np.zeros((nt, 6), dtype=np.int32), np.zeros((nt, 6), dtype=np.int8)
    re_, ro = np.zeros((nk, 3), dtype=np.int32), np.zeros((nk, 3), dtype=np.int8)
    cap = 6 * nt + 3 * nk
    edges = np.zeros((max(cap, 1), 2), dtype=np.int64)
    m = C.c_int64()
    ctx.check(ctx.lib.efb_build_edges(ctx.h, nt, _p(tc, i64p), nk, _p(rc_, i64p), _p(te, i32p), _p(to, i8p), _p(re_, i32p), _p(ro, i8p),
                                      C.byref(m), _p(edges, i64p), cap), "efb_build_edges")
    return te, to, re_, ro, edges[: m.value].copy()


class DeviceMesh:
    def __init__(self, ctx: Ctx, xyz, tet_nodes, tet_edges, tet_orient, tet_phys, edge_nodes):
        self.ctx = ctx
        self._keep = dict(
            xyz=np.ascontiguousarray(np.asarray(xyz, dtype=np.float64).reshape(-1, 3)),
            tet_nodes=_i32(tet_nodes).reshape(-1, 4), tet_edges=_i32(tet_edges).reshape(-1, 6),
            tet_orient=np.ascontiguousarray(np.asarray(tet_orient, dtype=np.int8).reshape(-1, 6)),
            tet_phys=_i32(tet_phys).reshape(-1), edge_nodes=_i32(edge_nodes).reshape(-1, 2))
        k = self._keep
        d = MeshDesc(k["xyz"].shape[0], _p(k["xyz"], f64p), k["tet_nodes"].shape[0], _p(k["tet_nodes"], i32p),
                     _p(k["tet_edges"], i32p), _p(k["tet_orient"], i8p), _p(k["tet_phys"], i32p), k["edge_nodes"].shape[0],
                     _p(k["edge_nodes"], i32p))
        h = C.c_void_p()
        ctx.check(ctx.lib.efb_mesh_create(ctx.h, C.byref(d), C.byref(h)), "efb_mesh_create")
        self.h = h
        self.n_edge = k["edge_nodes"].shape[0]
        self.n_tet = k["tet_nodes"].shape[0]
        self.n_node = k["xyz"].shape[0]
        n = ctx.lib.efb_mesh_num_slots(h)
        tags = np.zeros(n, dtype=np.int32)
        ctx.check(ctx.lib.efb_mesh_get_slot_tags(h, _p(tags, i32p)), "efb_mesh_get_slot_tags")
        self.slot_tags = tags

    def close(self):
        if self.h:
            self.ctx.lib.efb_mesh_destroy(self.h)
            self.h = None


def device_mesh_from_conn(ctx: "Ctx", xyz, tet_conn, tet_phys, tri_conn=None, node_ids=None):
    """Large-mesh ingest without the per-element host Mesh: number the edges on the GPU (efb_build_edges) and upload.
    tet_conn / tri_conn hold node IDS; node_ids (default 1..n) gives the id of every xyz row.
    Returns (DeviceMesh, info) with info = dict(tet_edges, tet_orient, tri_edges, tri_orient, edges (ids), edge_nodes (indices))."""
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype=np.float64).reshape(-1, 3))
    tc = np.asarray(tet_conn, dtype=np.int64).reshape(-1, 4)
    te, to, re_, ro, edges = build_edges_device(ctx, tc, tri_conn)
    if node_ids is None:
        tet_nodes, edge_nodes = (tc - 1).astype(np.int32), (edges - 1).astype(np.int32)
    else:
        ids = np.asarray(node_ids, dtype=np.int64)
        lut = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
        lut[ids] = np.arange(ids.size)
        tet_nodes, edge_nodes = lut[tc].astype(np.int32), lut[edges].astype(np.int32)
    dm = DeviceMesh(ctx, xyz, tet_nodes, te, to, tet_phys, edge_nodes)
    return dm, dict(tet_edges=te, tet_orient=to, tri_edges=re_, tri_orient=ro, edges=edges, edge_nodes=edge_nodes)


def pec_flags_from_tris(n_edges: int, tri_edges, tri_phys, pec_tag: int):
    """Dirichlet flags = edges of the boundary triangles tagged pec_tag (build_edge_pec, src/bc.cpp:47-80)."""
    flags = np.zeros(n_edges, dtype=np.uint8)
    sel = np.asarray(tri_phys) == pec_tag
    flags[np.asarray(tri_edges)[sel].reshape(-1)] = 1
    return flags


def make_materials(n_slots: int, eps=None, mu=None, eps_models=None, poles=None, pml=None):
    """Build an efb_materials struct; returns (struct, keepalive tuple)."""
    eps_a = _c128(np.ones(n_slots) if eps is None else eps)
    mu_a = _c128(np.ones(n_slots) if mu is None else mu)
    em = (Model * n_slots)()
    if eps_models is not None:
        for s, mdl in enumerate(eps_models):
            if mdl is not None:
                em[s] = mdl
    np_ = 0 if poles is None else len(poles)
    pa = (Pole * max(1, np_))()
    for i in range(np_):
        pa[i] = Pole(*poles[i])
    pm = (Pml * n_slots)()
    if pml is not None:
        for s, q in enumerate(pml):
            if q is not None:
                pm[s] = q
    m = Materials(n_slots, _p(eps_a.view(np.float64), f64p), _p(mu_a.view(np.float64), f64p), em, None, np_, pa, pm)
    return m, (eps_a, mu_a, em, pa, pm)


class DeviceSystem:
    def __init__(self, ctx: Ctx, h, mesh: Optional[DeviceMesh]):
        self.ctx, self.h, self.mesh = ctx, h, mesh
        m, nnz, nm, nr = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        ctx.check(ctx.lib.efb_system_dims(h, C.byref(m), C.byref(nnz), C.byref(nm), C.byref(nr)), "efb_system_dims")
        self.m, self.nnz, self.n_matrix, self.n_rhs = m.value, nnz.value, nm.value, nr.value

    @classmethod
    def from_mesh(cls, mesh: DeviceMesh, extra_rows=(), extra_cols=(), n_matrix=1, n_rhs=1):
        er, ec = _i32(extra_rows), _i32(extra_cols)
        h = C.c_void_p()
        ctx = mesh.ctx
        ctx.check(ctx.lib.efb_system_create(mesh.h, er.size, _p(er, i32p), _p(ec, i32p), n_matrix, n_rhs, C.byref(h)), "efb_system_create")
        return cls(ctx, h, mesh)

    @classmethod
    def from_mesh_rows(cls, mesh: DeviceMesh, row_begin: int, row_end: int, n_matrix=1, n_rhs=1):
        """Row block [row_begin, row_end) of the mesh's system (row-partitioned multi-GPU solve); `m` is the local row count."""
        h = C.c_void_p()
        ctx = mesh.ctx
        ctx.check(ctx.lib.efb_system_create_rows(mesh.h, row_begin, row_end, n_matrix, n_rhs, C.byref(h)), "efb_system_create_rows")
        s = cls(ctx, h, mesh)
        s.row_begin, s.row_end = row_begin, row_end
        return s

    @classmethod
    def from_csr(cls, ctx: Ctx, rowptr, colidx, vals=None, n_matrix=1, n_rhs=1):
        rp, ci = _i32(rowptr), _i32(colidx)
        v = None if vals is None else _c128(vals)
        h = C.c_void_p()
        ctx.check(ctx.lib.efb_system_create_csr(ctx.h, rp.size - 1, ci.size, _p(rp, i32p), _p(ci, i32p),
                                                None if v is None else _p(v.view(np.float64), f64p), n_matrix, n_rhs, C.byref(h)),
                  "efb_system_create_csr")
        return cls(ctx, h, None)

    def pattern(self):
        rp = np.zeros(self.m + 1, dtype=np.int32)
        ci = np.zeros(self.nnz, dtype=np.int32)
        self.ctx.check(self.ctx.lib.efb_system_get_pattern(self.h, _p(rp, i32p), _p(ci, i32p)), "efb_system_get_pattern")
        return rp, ci

    def values(self, matrix=0) -> np.ndarray:
        v = np.zeros(self.nnz, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.efb_system_get_values(self.h, matrix, _p(v.view(np.float64), f64p)), "efb_system_get_values")
        return v

    def set_values(self, matrix, vals):
        v = _c128(vals)
        assert v.size == self.nnz
        self.ctx.check(self.ctx.lib.efb_system_set_values(self.h, matrix, _p(v.view(np.float64), f64p)), "efb_system_set_values")

    def set_dirichlet(self, flags):
        f = np.ascontiguousarray(np.asarray(flags, dtype=np.uint8))
        assert f.size == (self.mesh.n_edge if getattr(self, "row_begin", None) is not None else self.m)  # row blocks take GLOBAL flags
        self.ctx.check(self.ctx.lib.efb_system_set_dirichlet(self.h, _p(f, u8p)), "efb_system_set_dirichlet")

    def set_gradient(self, n_node, edge_nodes):
        en = _i32(edge_nodes).reshape(-1, 2)
        self.ctx.check(self.ctx.lib.efb_system_set_gradient(self.h, n_node, _p(en, i32p)), "efb_system_set_gradient")

    def assemble_volume(self, omegas, materials: Materials, first=0, mode=0):
        om = np.ascontiguousarray(np.asarray(omegas, dtype=np.float64).reshape(-1))
        self.ctx.check(self.ctx.lib.efb_assemble_volume(self.h, first, om.size, _p(om, f64p), C.byref(materials), mode), "efb_assemble_volume")

    def combine_km(self, dst_first, k0sq, src_k, src_m):
        q = np.ascontiguousarray(np.asarray(k0sq, dtype=np.float64).reshape(-1))
        self.ctx.check(self.ctx.lib.efb_combine_km(self.h, dst_first, q.size, _p(q, f64p), src_k, src_m), "efb_combine_km")

    def add_diag(self, edges, coef, first=0):
        e = _i32(edges)
        cf = _c128(coef).reshape(-1)
        self.ctx.check(self.ctx.lib.efb_add_diag(self.h, first, cf.size, e.size, _p(e, i32p), _p(cf.view(np.float64), f64p)), "efb_add_diag")

    def rhs_set(self, rhs, b):
        v = _c128(b)
        self.ctx.check(self.ctx.lib.efb_rhs_set(self.h, rhs, _p(v.view(np.float64), f64p)), "efb_rhs_set")

    def rhs_zero(self, rhs):
        self.ctx.check(self.ctx.lib.efb_rhs_zero(self.h, rhs), "efb_rhs_zero")

    def rhs_get(self, rhs) -> np.ndarray:
        v = np.zeros(self.m, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.efb_rhs_get(self.h, rhs, _p(v.view(np.float64), f64p)), "efb_rhs_get")
        return v

    def x_get(self, rhs) -> np.ndarray:
        v = np.zeros(self.m, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.efb_x_get(self.h, rhs, _p(v.view(np.float64), f64p)), "efb_x_get")
        return v

    def x_set(self, rhs, x):
        v = _c128(x)
        self.ctx.check(self.ctx.lib.efb_x_set(self.h, rhs, _p(v.view(np.float64), f64p)), "efb_x_set")

    def x_recover(self, rhs, dst, src, phase):
        d, s, ph = _i32(dst), _i32(src), _c128(phase)
        self.ctx.check(self.ctx.lib.efb_x_recover(self.h, rhs, d.size, _p(d, i32p), _p(s, i32p), _p(ph.view(np.float64), f64p)), "efb_x_recover")

    def apply_periodic(self, master, slave, phase, first=0, count=1):
        ms, sl, ph = _i32(master), _i32(slave), _c128(phase)
        self.ctx.check(self.ctx.lib.efb_apply_periodic(self.h, first, count, ms.size, _p(ms, i32p), _p(sl, i32p), _p(ph.view(np.float64), f64p)), "efb_apply_periodic")

    def solve(self, first=0, count=None, method=METHOD_AUTO, precond=PRECOND_AUX, tol=1e-10, max_iterations=10000,
              check_every=0, symmetric=True, zero_initial_guess=True, max_restarts=3):
        count = self.n_matrix - first if count is None else count
        o = SolveOpts(method, precond, tol, max_iterations, check_every, 1 if symmetric else 0, 1 if zero_initial_guess else 0, max_restarts, 0)
        res = (SolveResult * (count * self.n_rhs))()
        self.ctx.check(self.ctx.lib.efb_solve(self.h, first, count, C.byref(o), res), "efb_solve")
        return [dict(iters=r.iters, converged=bool(r.converged), method=r.method, precond=r.precond, residual=r.residual) for r in res]

    def solve_direct(self, first=0, count=None):
        """Dense LU with partial pivoting on the device (efb_solve_direct): the robust fallback for small systems."""
        count = self.n_matrix - first if count is None else count
        res = (SolveResult * (count * self.n_rhs))()
        self.ctx.check(self.ctx.lib.efb_solve_direct(self.h, first, count, res), "efb_solve_direct")
        return [dict(iters=r.iters, converged=bool(r.converged), method=r.method, precond=r.precond, residual=r.residual) for r in res]

    def last_solve_kernel_ms(self) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.lib.efb_system_last_solve_kernel_ms(self.h, C.byref(ms)), "efb_system_last_solve_kernel_ms")
        return ms.value

    def last_solve_shape(self):
        """(CTAs per cluster, rhs per job, resident clusters) of the last solve; (0, 0, 0) if it did not use the cluster kernel."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self.ctx.check(self.ctx.lib.efb_system_last_solve_shape(self.h, C.byref(a), C.byref(b), C.byref(c)), "efb_system_last_solve_shape")
        return a.value, b.value, c.value

    def spmv(self, matrix, x) -> np.ndarray:
        xv = _c128(x)
        y = np.zeros(self.m, dtype=np.complex128)
        self.ctx.check(self.ctx.lib.efb_spmv_host(self.h, matrix, _p(xv.view(np.float64), f64p), _p(y.view(np.float64), f64p)), "efb_spmv_host")
        return y

    def dist_solve(self, tol=1e-10, max_iterations=10000, precond=PRECOND_JACOBI, check_every=0, zero_initial_guess=True, max_restarts=3,
                   halo_mode=0):
        o = SolveOpts(METHOD_COCG, precond, tol, max_iterations, check_every, 1, 1 if zero_initial_guess else 0, max_restarts, 0)
        r = SolveResult()
        self.ctx.check(self.ctx.lib.efb_dist_solve(self.h, C.byref(o), C.byref(r), halo_mode), "efb_dist_solve")
        return dict(iters=r.iters, converged=bool(r.converged), method=r.method, precond=r.precond, residual=r.residual)

    def dist_bench(self, which, reps, halo_mode=0) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.lib.efb_dist_bench(self.h, which, reps, halo_mode, C.byref(ms)), "efb_dist_bench")
        return ms.value

    def huygens_eval(self, rhs, tri_nodes, tri_tet, tri_tet_edges, omega, mu_r=1.0):
        """Tangential E, H at the centroids of surface triangles from solution `rhs` (efb_huygens_eval).
        Returns dict(r [n,3], n [n,3], E_tan [n,3] c128, H_tan [n,3] c128, area [n])."""
        tn, tt, te = _i32(tri_nodes).reshape(-1, 3), _i32(tri_tet).reshape(-1), _i32(tri_tet_edges).reshape(-1, 6)
        n = tt.size
        r, nn = np.zeros((n, 3)), np.zeros((n, 3))
        E, H = np.zeros((n, 3), dtype=np.complex128), np.zeros((n, 3), dtype=np.complex128)
        area = np.zeros(n)
        mu = np.array([complex(mu_r)], dtype=np.complex128)
        self.ctx.check(self.ctx.lib.efb_huygens_eval(self.mesh.h, self.h, rhs, None, n, _p(tn, i32p), _p(tt, i32p), _p(te, i32p), float(omega), _p(mu.view(np.float64), f64p),
                                                     _p(r, f64p), _p(nn, f64p), _p(E.view(np.float64), f64p), _p(H.view(np.float64), f64p),
                                                     _p(area, f64p)), "efb_huygens_eval")
        return dict(r=r, n=nn, E_tan=E, H_tan=H, area=area)

    def bench_kernel(self, which, reps) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.lib.efb_bench_kernel(self.h, which, reps, C.byref(ms)), "efb_bench_kernel")
        return ms.value

    def close(self):
        if self.h:
            self.ctx.lib.efb_system_destroy(self.h)
            self.h = None


class DevicePort:
    def __init__(self, sys: DeviceSystem, edges, weights, ms_rows=(), ms_cols=(), ms_vals=()):
        self.sys = sys
        e, w = _i32(edges), _c128(weights)
        r, c_ = _i32(ms_rows), _i32(ms_cols)
        v = np.ascontiguousarray(np.asarray(ms_vals, dtype=np.float64))
        h = C.c_void_p()
        ctx = sys.ctx
        ctx.check(ctx.lib.efb_port_create(sys.h, e.size, _p(e, i32p), _p(w.view(np.float64), f64p), r.size, _p(r, i32p), _p(c_, i32p),
                                          _p(v, f64p), C.byref(h)), "efb_port_create")
        self.h = h

    def _coef(self, coef):
        return _c128(coef).reshape(-1)

    def normalize_mass(self) -> float:
        ns = C.c_double()
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_normalize_mass(self.h, C.byref(ns)), "efb_port_normalize_mass")
        return ns.value

    def add_block(self, coef, first=0):
        cf = self._coef(coef)
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_add_block(self.sys.h, self.h, first, cf.size, _p(cf.view(np.float64), f64p)), "efb_port_add_block")

    def add_mass(self, coef, first=0):
        cf = self._coef(coef)
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_add_mass(self.sys.h, self.h, first, cf.size, _p(cf.view(np.float64), f64p)), "efb_port_add_mass")

    def rhs_weights(self, rhs, coef):
        cf = self._coef([coef])
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_rhs_weights(self.sys.h, self.h, rhs, _p(cf.view(np.float64), f64p)), "efb_port_rhs_weights")

    def rhs_mass(self, rhs, coef):
        cf = self._coef([coef])
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_rhs_mass(self.sys.h, self.h, rhs, _p(cf.view(np.float64), f64p)), "efb_port_rhs_mass")

    def rhs_batch(self, rhs_idx, coef, use_mass: bool):
        idx, cf = _i32(rhs_idx), self._coef(coef)
        assert idx.size == cf.size
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_rhs_batch(self.sys.h, self.h, idx.size, _p(idx, i32p), _p(cf.view(np.float64), f64p),
                                                               1 if use_mass else 0), "efb_port_rhs_batch")

    def project_batch(self, rhs_idx, use_mass: bool) -> np.ndarray:
        idx = _i32(rhs_idx)
        out = np.zeros(idx.size, dtype=np.complex128)
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_project_batch(self.sys.h, self.h, idx.size, _p(idx, i32p), _p(out.view(np.float64), f64p),
                                                                   1 if use_mass else 0), "efb_port_project_batch")
        return out

    def project_weights(self, rhs) -> complex:
        out = np.zeros(1, dtype=np.complex128)
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_project_weights(self.sys.h, self.h, rhs, _p(out.view(np.float64), f64p)), "efb_port_project_weights")
        return complex(out[0])

    def project_mass(self, rhs) -> complex:
        out = np.zeros(1, dtype=np.complex128)
        self.sys.ctx.check(self.sys.ctx.lib.efb_port_project_mass(self.sys.h, self.h, rhs, _p(out.view(np.float64), f64p)), "efb_port_project_mass")
        return complex(out[0])

    def close(self):
        if self.h:
            self.sys.ctx.lib.efb_port_destroy(self.h)
            self.h = None
