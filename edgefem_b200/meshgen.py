"""Deterministic structured tetrahedral mesh generators + a Gmsh v2 ASCII writer (product-side
utilities; gmsh itself is not available in this environment).

All generators return plain numpy arrays ``(xyz, tets, tet_phys, tris, tri_phys)`` with 1-based
node ids implied (node id = index + 1), ready for ``pyedgefem.mesh_from_arrays``.
Every box is split into 6 tetrahedra with the Kuhn (Freudenthal) triangulation, which is
face-compatible between neighbouring boxes and translation invariant -- opposite faces of a
block carry identical triangulations, so periodic face pairs are node-matched by construction.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np

# the 6 Kuhn tets of the unit cube: paths 000 -> 111 adding one axis at a time
_KUHN = []
for perm in itertools.permutations(range(3)):
    v = [np.zeros(3, dtype=np.int64)]
    for ax in perm:
        nxt = v[-1].copy()
        nxt[ax] = 1
        v.append(nxt)
    _KUHN.append(np.array(v))
_KUHN = np.array(_KUHN)  # [6,4,3]


def box_grid(xs: np.ndarray, ys: np.ndarray, zs: np.ndarray):
    """Tensor-product grid -> (xyz [n,3], tets [6*nc,4] 1-based ids, cell index of every tet [6*nc,3])."""
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k + 1

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    tets = np.empty((I.size, 6, 4), dtype=np.int64)
    for t in range(6):
        for c in range(4):
            o = _KUHN[t, c]
            tets[:, t, c] = nid(I + o[0], J + o[1], K + o[2])
    cells = np.repeat(np.stack([I, J, K], axis=1)[:, None, :], 6, axis=1).reshape(-1, 3)
    return xyz, tets.reshape(-1, 4), cells


def boundary_faces(tets: np.ndarray):
    """All faces that belong to exactly one tet: (tris [k,3] node ids, owning tet index [k])."""
    f = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]], axis=0)
    owner = np.tile(np.arange(tets.shape[0]), 4)
    key = np.sort(f, axis=1)
    uniq, idx, cnt = np.unique(key, axis=0, return_index=True, return_counts=True)
    sel = idx[cnt == 1]
    sel.sort()
    return f[sel], owner[sel]


def grid_boundary_faces(tets: np.ndarray, nx: int, ny: int, nz: int):
    """Boundary faces of a box_grid() mesh without sorting: a face is on the boundary iff its three
    nodes share a boundary grid plane.  Returns (tris [k,3], plane [k] in 0..5 = x0,x1,y0,y1,z0,z1).
    O(N_tet) time and memory -- usable for the 20 M-tet cube where np.unique over faces is not."""
    out_t, out_p = [], []
    for cols in ([0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]):
        f = tets[:, cols]
        z0 = f - 1
        k = z0 % (nz + 1)
        j = (z0 // (nz + 1)) % (ny + 1)
        i = z0 // ((nz + 1) * (ny + 1))
        for pid, (c, v) in enumerate(((i, 0), (i, nx), (j, 0), (j, ny), (k, 0), (k, nz))):
            sel = np.all(c == v, axis=1)
            if sel.any():
                out_t.append(f[sel])
                out_p.append(np.full(int(sel.sum()), pid, dtype=np.int32))
    return np.concatenate(out_t, axis=0), np.concatenate(out_p)


def faces_on_plane(xyz: np.ndarray, tris: np.ndarray, axis: int, value: float, tol: float = 1e-12) -> np.ndarray:
    c = xyz[tris - 1][:, :, axis]
    return np.all(np.abs(c - value) <= tol, axis=1)


def cube_cavity(n: int, length: float = 1.0, jitter: float = 0.0, seed: int = 1234, pec_tag: int = 1, vol_tag: int = 100):
    """Unit cube, n^3 boxes x 6 Kuhn tets, every boundary face tagged PEC (SURVEY config C5).
    ``jitter`` perturbs interior nodes by +-jitter*h (seeded) to avoid structured-mesh luck."""
    g = np.linspace(0.0, length, n + 1)
    xyz, tets, _ = box_grid(g, g, g)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        h = length / n
        interior = np.all((xyz > 1e-12) & (xyz < length - 1e-12), axis=1)
        xyz = xyz.copy()
        xyz[interior] += (rng.random((int(interior.sum()), 3)) * 2.0 - 1.0) * jitter * h
    tris, _ = grid_boundary_faces(tets, n, n, n)
    return xyz, tets, np.full(tets.shape[0], vol_tag, dtype=np.int32), tris, np.full(tris.shape[0], pec_tag, dtype=np.int32)


def rect_waveguide(a: float = 0.02286, b: float = 0.01016, length: float = 0.05, nx: int = 8, ny: int = 4, nz: int = 18,
                   pec_tag: int = 1, port1_tag: int = 2, port2_tag: int = 3, vol_tag: int = 100):
    """Structured WR-90-like guide along z: walls PEC, z=0 -> port 1, z=L -> port 2."""
    xyz, tets, _ = box_grid(np.linspace(0, a, nx + 1), np.linspace(0, b, ny + 1), np.linspace(0, length, nz + 1))
    tris, _ = boundary_faces(tets)
    tags = np.full(tris.shape[0], pec_tag, dtype=np.int32)
    tags[faces_on_plane(xyz, tris, 2, 0.0)] = port1_tag
    tags[faces_on_plane(xyz, tris, 2, length)] = port2_tag
    return xyz, tets, np.full(tets.shape[0], vol_tag, dtype=np.int32), tris, tags


def write_gmsh_v2(path: str, xyz, tets, tet_phys, tris, tri_phys, node_ids=None) -> None:
    """Gmsh 2.2 ASCII: triangles (type 2) first, then tetrahedra (type 4), two tags each."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    ids = np.arange(1, xyz.shape[0] + 1) if node_ids is None else np.asarray(node_ids)
    tets = np.asarray(tets).reshape(-1, 4)
    tris = np.asarray(tris).reshape(-1, 3)
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % xyz.shape[0])
        for i in range(xyz.shape[0]):
            f.write("%d %.17g %.17g %.17g\n" % (ids[i], xyz[i, 0], xyz[i, 1], xyz[i, 2]))
        f.write("$EndNodes\n$Elements\n%d\n" % (tets.shape[0] + tris.shape[0]))
        eid = 1
        for k in range(tris.shape[0]):
            f.write("%d 2 2 %d %d %d %d %d\n" % (eid, tri_phys[k], tri_phys[k], tris[k, 0], tris[k, 1], tris[k, 2]))
            eid += 1
        for k in range(tets.shape[0]):
            f.write("%d 4 2 %d %d %d %d %d %d\n" % (eid, tet_phys[k], tet_phys[k], tets[k, 0], tets[k, 1], tets[k, 2], tets[k, 3]))
            eid += 1
        f.write("$EndElements\n")


def unit_cell(px: float = 0.005, py: float = 0.005, h_sub: float = 0.0005, h_air: float = 0.004, nx: int = 4, ny: int = 4,
              nz_sub: int = 1, nz_air: int = 4, patch: Optional[Tuple[float, float]] = None, tags: Optional[Dict[str, int]] = None):
    """Periodic unit cell (tag scheme of the reference's unit_cell design, python/edgefem/designs/unit_cell.py:62-72):
    ground plane z=0 -> PEC(1); optional centred patch on the substrate top -> 2 (interior faces); top z=H -> port(4);
    x=0 / x=px -> periodic master/slave (5/6); y=0 / y=py -> (7/8); substrate volume 100, air 101.
    Opposite faces are node-matched by construction (translation-invariant Kuhn split)."""
    t = dict(pec=1, patch=2, port=4, xm=5, xs=6, ym=7, ys=8, sub=100, air=101)
    if tags:
        t.update(tags)
    zs = np.concatenate([np.linspace(0, h_sub, nz_sub + 1), np.linspace(h_sub, h_sub + h_air, nz_air + 1)[1:]])
    xyz, tets, cells = box_grid(np.linspace(0, px, nx + 1), np.linspace(0, py, ny + 1), zs)
    tet_phys = np.where(cells[:, 2] < nz_sub, t["sub"], t["air"]).astype(np.int32)
    tris, _ = boundary_faces(tets)
    H = h_sub + h_air
    tri_phys = np.zeros(tris.shape[0], dtype=np.int32)
    tri_phys[faces_on_plane(xyz, tris, 2, 0.0)] = t["pec"]
    tri_phys[faces_on_plane(xyz, tris, 2, H, tol=1e-12)] = t["port"]
    tri_phys[faces_on_plane(xyz, tris, 0, 0.0)] = t["xm"]
    tri_phys[faces_on_plane(xyz, tris, 0, px)] = t["xs"]
    tri_phys[faces_on_plane(xyz, tris, 1, 0.0)] = t["ym"]
    tri_phys[faces_on_plane(xyz, tris, 1, py)] = t["ys"]
    if patch is not None:
        # interior faces on z = h_sub inside the centred patch rectangle
        f = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]], axis=0)
        key = np.sort(f, axis=1)
        _, idx = np.unique(key, axis=0, return_index=True)
        f = f[np.sort(idx)]
        c = xyz[f - 1]
        on = np.all(np.abs(c[:, :, 2] - h_sub) < 1e-12, axis=1)
        cx, cy = c[:, :, 0].mean(axis=1), c[:, :, 1].mean(axis=1)
        inside = on & (np.abs(cx - px / 2) < patch[0] / 2) & (np.abs(cy - py / 2) < patch[1] / 2)
        tris = np.concatenate([tris, f[inside]], axis=0)
        tri_phys = np.concatenate([tri_phys, np.full(int(inside.sum()), t["patch"], dtype=np.int32)])
    return xyz, tets, tet_phys, tris, tri_phys


def _lines(breaks: Sequence[float], hmax: float) -> np.ndarray:
    """Grid lines through every break point with uniform spacing <= hmax inside each interval."""
    b = sorted(set(float(round(x, 12)) for x in breaks))
    out = [b[0]]
    for lo, hi in zip(b[:-1], b[1:]):
        n = max(1, int(math.ceil((hi - lo) / hmax - 1e-9)))
        out.extend(lo + (hi - lo) * (k + 1) / n for k in range(n))
    return np.array(out)


def _unique_faces(tets: np.ndarray) -> np.ndarray:
    """Every face of the mesh once (first-seen orientation), boundary and interior."""
    f = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]], axis=0)
    key = np.sort(f, axis=1)
    _, idx = np.unique(key, axis=0, return_index=True)
    return f[np.sort(idx)]


def patch_antenna(patch_length: float = 28.6e-3, patch_width: float = 37.3e-3, substrate_height: float = 1.6e-3, probe_y_offset: float = -9.0e-3,
                  probe_x_offset: float = 0.0, design_freq: float = 2.45e9, density: float = 12.0, ground_plane_size: Optional[float] = None,
                  air_box_height: Optional[float] = None, cavity_depth: float = 5e-3, probe_radius: float = 0.5e-3, nz_sub: int = 2,
                  nz_cavity: int = 2, hmax_scale: float = 1.0, huygens: bool = True):
    """Probe-fed patch antenna on a grounded substrate (BASELINE config 3), structured Kuhn tets with the geometry and
    the physical tags of the reference's gmsh model (python/edgefem/designs/stacked_patch.py:63-73,184-330 as driven by
    patch_antenna.py:213-243 and examples/run_patch_fullwave.py:19-29; gmsh itself is not available here):
    domain [0,gp]^2 x [-cavity_depth, h + h_air], gp = 3 x max(patch side), h_air = lambda/4 at the design frequency;
    ground plane z=0 -> 1 (minus the square probe hole of half-width pw), cavity walls (sides below ground) -> 2, cavity
    bottom -> 3, probe conductor (square column of half-width pr from the bottom to the patch) side faces -> 4, lumped
    port strip [cx+pr, cx+pw] x [cy-pr, cy+pr] at z=0 -> 5, patch (width along x, length along y, centred) at z=h -> 10,
    top + sides above ground -> 50 (ABC), Huygens surface at mid air height, inset 0.1 gp -> 60; volumes: cavity 100,
    substrate 110, air 150.  pr = max(probe_radius, lc/8, 0.5 mm), pw = max(2.5 pr, lc/4, 2 mm), lc = lambda/density
    (stacked_patch.py:190-195).  Returns (xyz, tets, tet_phys, tris, tri_phys, info)."""
    c0 = 299792458.0
    lam = c0 / design_freq
    lc = lam / density
    gp = ground_plane_size or 3.0 * max(patch_length, patch_width)
    h_air = air_box_height or lam / 4.0
    h = substrate_height
    pr = max(probe_radius, lc / 8.0, 0.5e-3)
    pw = max(2.5 * pr, lc / 4.0, 2e-3)
    cx, cy = gp / 2 + probe_x_offset, gp / 2 + probe_y_offset
    px0, px1 = gp / 2 - patch_width / 2, gp / 2 + patch_width / 2
    py0, py1 = gp / 2 - patch_length / 2, gp / 2 + patch_length / 2
    z_bot, z_top = -cavity_depth, h + h_air
    z_huy = h + 0.5 * h_air
    inset = 0.1 * gp
    hm = lc * hmax_scale
    xs = _lines([0.0, gp, px0, px1, cx - pw, cx - pr, cx + pr, cx + pw] + ([inset, gp - inset] if huygens else []), hm)
    ys = _lines([0.0, gp, py0, py1, cy - pw, cy - pr, cy + pr, cy + pw] + ([inset, gp - inset] if huygens else []), hm)
    zs = np.concatenate([np.linspace(z_bot, 0.0, nz_cavity + 1), np.linspace(0.0, h, nz_sub + 1)[1:],
                         _lines([h, z_huy, z_top] if huygens else [h, z_top], hm)[1:]])
    xyz, tets, cells = box_grid(xs, ys, zs)
    zc = 0.5 * (zs[cells[:, 2]] + zs[cells[:, 2] + 1])
    tet_phys = np.where(zc < 0.0, 100, np.where(zc < h, 110, 150)).astype(np.int32)
    faces = _unique_faces(tets)
    P = xyz[faces - 1]                      # [k,3 nodes,3]
    cen = P.mean(axis=1)
    tol = 1e-9

    def on_plane(axis, v):
        return np.all(np.abs(P[:, :, axis] - v) < tol, axis=1)

    def inside(x0, x1, y0, y1):
        return (cen[:, 0] > x0) & (cen[:, 0] < x1) & (cen[:, 1] > y0) & (cen[:, 1] < y1)

    tag = np.zeros(faces.shape[0], dtype=np.int32)
    z0 = on_plane(2, 0.0)
    hole = inside(cx - pw, cx + pw, cy - pw, cy + pw)
    tag[z0 & ~hole] = 1
    tag[z0 & inside(cx + pr, cx + pw, cy - pr, cy + pr)] = 5
    side = on_plane(0, 0.0) | on_plane(0, gp) | on_plane(1, 0.0) | on_plane(1, gp)
    tag[side & (cen[:, 2] < 0.0)] = 2
    tag[on_plane(2, z_bot)] = 3
    col_z = (cen[:, 2] > z_bot) & (cen[:, 2] < h)
    in_y = (cen[:, 1] > cy - pr) & (cen[:, 1] < cy + pr)
    in_x = (cen[:, 0] > cx - pr) & (cen[:, 0] < cx + pr)
    tag[(on_plane(0, cx - pr) | on_plane(0, cx + pr)) & in_y & col_z] = 4
    tag[(on_plane(1, cy - pr) | on_plane(1, cy + pr)) & in_x & col_z] = 4
    tag[on_plane(2, h) & inside(px0, px1, py0, py1)] = 10
    tag[on_plane(2, z_top)] = 50
    tag[side & (cen[:, 2] > 0.0)] = 50
    if huygens:
        tag[on_plane(2, z_huy) & inside(inset, gp - inset, inset, gp - inset)] = 60
    keep = tag != 0
    info = dict(gp=gp, h_air=h_air, lc=lc, pr=pr, pw=pw, cx=cx, cy=cy, grid=(len(xs) - 1, len(ys) - 1, len(zs) - 1))
    return xyz, tets, tet_phys, faces[keep], tag[keep], info
