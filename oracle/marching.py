"""Oracle isosurface extraction: marching cubes over the sign-change cells of an SDF slab.

Reference: none (`/root/reference/README.md:1`); this is SURVEY.md section 8(f) row N1, the consumer
of the sign-change mask.  Case tables: oracle/mc_tables.py (data written by tools/gen_mc_tables.py).  Test infrastructure only.

Definition (what the CUDA kernels in csrc/marching.cu must reproduce bit for bit):
  * cells in C order (z, y, x), x fastest; a cell's triangles in table order; output = triangle soup
    [n_tri, 3 vertices, 3 coords (x, y, z)] float32;
  * vertex on edge (lo, hi), lo < hi along the edge's axis:  t = v_lo / (v_lo - v_hi);
    p = c_lo + t * (c_hi - c_lo) on that axis (each operation rounded to float32, no fused multiply-add),
    the other two coordinates are the node coordinates of A1;
  * normals point from inside (v < 0) to outside.
"""
from __future__ import annotations

import numpy as np

from .grid import axis_coords
from .mc_tables import CORNERS, EDGES, EDGE_AXIS, MC_NTRI, MC_TRI


def marching_cubes(sdf: np.ndarray, res: int | None = None, z0: int = 0) -> np.ndarray:
    s = np.ascontiguousarray(sdf, dtype=np.float32)
    nz, ny, nx = s.shape
    res = nx if res is None else res
    cz, cy, cx = nz - 1, ny - 1, nx - 1
    if min(cz, cy, cx) < 1:
        return np.zeros((0, 3, 3), np.float32)
    coords = axis_coords(res)
    inside = s < 0
    case = np.zeros((cz, cy, cx), dtype=np.int64)
    for i, (dx, dy, dz) in enumerate(CORNERS):
        case |= inside[dz:dz + cz, dy:dy + cy, dx:dx + cx].astype(np.int64) << i
    case = case.ravel()
    ntri = MC_NTRI[case].astype(np.int64)
    first = np.concatenate([[0], np.cumsum(ntri)])
    out = np.zeros((int(first[-1]), 3, 3), dtype=np.float32)
    active = np.nonzero(ntri)[0]
    if active.size == 0:
        return out
    ix = active % cx
    iy = (active // cx) % cy
    iz = active // (cx * cy)
    for slot in range(MC_TRI.shape[1]):
        sel = ntri[active] * 3 > slot
        if not sel.any():
            break
        cells = active[sel]
        e = MC_TRI[case[cells], slot]
        lo = EDGES[e, 0]
        axis = EDGE_AXIS[e]
        x0 = ix[sel] + CORNERS[lo, 0]
        y0 = iy[sel] + CORNERS[lo, 1]
        z0l = iz[sel] + CORNERS[lo, 2]
        x1 = x0 + (axis == 0)
        y1 = y0 + (axis == 1)
        z1l = z0l + (axis == 2)
        v_lo = s[z0l, y0, x0]
        v_hi = s[z1l, y1, x1]
        t = (v_lo / (v_lo - v_hi)).astype(np.float32)
        p = np.stack([coords[x0], coords[y0], coords[z0 + z0l]], axis=1).astype(np.float32)
        c_lo = np.where(axis == 0, coords[x0], np.where(axis == 1, coords[y0], coords[z0 + z0l])).astype(np.float32)
        c_hi = np.where(axis == 0, coords[x1], np.where(axis == 1, coords[y1], coords[z0 + z1l])).astype(np.float32)
        d = (c_hi - c_lo).astype(np.float32)
        pv = (c_lo + (t * d).astype(np.float32)).astype(np.float32)
        p[np.arange(p.shape[0]), axis] = pv
        out[first[cells] + slot // 3, slot % 3] = p
    return out


def mesh_is_closed(tris: np.ndarray) -> bool:
    """Every directed edge of the soup is matched by exactly one opposite edge (watertight, consistently oriented)."""
    v = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 3)
    _, idx = np.unique(v.view(np.uint32).reshape(-1, 3), axis=0, return_inverse=True)
    f = idx.reshape(-1, 3)
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]     # zero-area slivers (t = 0 or 1) drop out
    a = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    fwd = a[:, 0].astype(np.int64) * (idx.max() + 1) + a[:, 1]
    bwd = a[:, 1].astype(np.int64) * (idx.max() + 1) + a[:, 0]
    uf, cf = np.unique(fwd, return_counts=True)
    ub, cb = np.unique(bwd, return_counts=True)
    return bool(np.array_equal(uf, ub) and np.array_equal(cf, cb))
