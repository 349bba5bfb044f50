"""ctypes view of oracle/oracle_assembly.c (test infrastructure; see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle_assembly.so")


def load():
    if not os.path.exists(_LIB):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = C.CDLL(_LIB)
    lib.efo_volume_triplets.restype = None
    return lib


def volume_matrix(mesh, omega, eps_tet, mu_tet):
    """Unmasked volume matrix sum_t (K/mu - k0^2 eps M) s s^T via the C restatement (CSR, duplicates summed)."""
    lib = load()
    nt = mesh.tet_conn.shape[0]
    xyz = np.ascontiguousarray(mesh.xyz, dtype=np.float64)
    tn = np.ascontiguousarray(mesh.node_idx_of(mesh.tet_conn), dtype=np.int32)
    te = np.ascontiguousarray(mesh.tet_edges, dtype=np.int32)
    to = np.ascontiguousarray(mesh.tet_orient, dtype=np.int32)
    eps = np.ascontiguousarray(np.broadcast_to(np.asarray(eps_tet, dtype=np.complex128), (nt,)))
    mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu_tet, dtype=np.complex128), (nt,)))
    rows = np.empty(36 * nt, dtype=np.int32)
    cols = np.empty(36 * nt, dtype=np.int32)
    vals = np.empty(36 * nt, dtype=np.complex128)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.efo_volume_triplets(C.c_int64(nt), p(xyz), p(tn), p(te), p(to), p(eps), p(mu), C.c_double(omega), p(rows), p(cols), p(vals))
    m = mesh.num_edges
    A = sp.coo_matrix((vals, (rows, cols)), shape=(m, m)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A
