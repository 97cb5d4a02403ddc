"""Marching-cubes case tables, GENERATED (no table was copied from anywhere; the mounted reference has
no code, `/root/reference/README.md:1`, and there is no network).

Construction, for each of the 256 corner-sign configurations:
  * a cube edge is *crossed* when its two corners differ in inside(v) := v < 0;
  * on every cube face the crossed edges (0, 2 or 4 of them) are joined by segments; on an ambiguous
    face (4 crossings, signs alternating) each INSIDE corner is cut off on its own - a rule that
    depends only on the face's four signs, so the two cubes sharing a face always agree and the mesh
    is crack-free;
  * every crossed edge then has exactly two segments: the segments form closed loops, each loop is
    oriented so that its normal points from the inside corners to the outside ones and is cut
    into a triangle fan.
Corner i sits at (x, y, z) = (i & 1, (i >> 1) & 1, i >> 2); case index = sum(inside(corner i) << i).
Edge e = 4 * axis + k joins corner `lo` (k-th corner, ascending, whose `axis` bit is 0) to lo | (1 << axis);
vertices are always interpolated from lo to hi, so the two to four cubes sharing an edge compute the
same floating-point expression.

Build tool of the product: `python tools/gen_mc_tables.py` writes
    latent-diffusion-models-for-shape-sdfs_b200/csrc/mc_tables.h   (what the GPU kernels include) and
    oracle/mc_tables.npz                                            (the same numbers as data for the CPU oracle);
a CPU test checks that both are up to date, and tests/test_marching.py checks the tables themselves against properties
that do not come from this construction (analytic shapes, orientation along the trilinear field, closedness).
"""
from __future__ import annotations

import os

import numpy as np

CORNERS = np.array([[i & 1, (i >> 1) & 1, i >> 2] for i in range(8)], dtype=np.int64)
EDGES = []
for axis in range(3):
    for c in range(8):
        if not (c >> axis) & 1:
            EDGES.append((c, c | (1 << axis)))
EDGES = np.array(EDGES, dtype=np.int64)                 # [12, 2] (lo, hi)
EDGE_AXIS = np.repeat(np.arange(3), 4)
_EDGE_ID = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGES)}


def _face_cycles():
    """The six faces as corner cycles (each consecutive pair is a cube edge)."""
    faces = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in (0, 1):
            base = side << axis
            faces.append([base, base | (1 << u), base | (1 << u) | (1 << v), base | (1 << v)])
    return faces


FACES = _face_cycles()


def _edge_between(a: int, b: int) -> int:
    return _EDGE_ID[(min(a, b), max(a, b))]


def _case_triangles(case: int):
    inside = [(case >> i) & 1 for i in range(8)]
    adj = {}                                             # crossed edge -> its two neighbours along the surface

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)

    for cyc in FACES:
        crossed = [k for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]   # face edge k = (cyc[k], cyc[k+1])
        if len(crossed) == 2:
            link(_edge_between(cyc[crossed[0]], cyc[(crossed[0] + 1) % 4]), _edge_between(cyc[crossed[1]], cyc[(crossed[1] + 1) % 4]))
        elif len(crossed) == 4:
            for k in range(4):                           # cut off every inside corner: join the two face edges that meet in it
                if inside[cyc[k]]:
                    link(_edge_between(cyc[k - 1], cyc[k]), _edge_between(cyc[k], cyc[(k + 1) % 4]))
    assert all(len(v) == 2 for v in adj.values()), (case, adj)
    mid = {e: (CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]) / 2.0 for e in adj}
    out_dir = {}
    for e in adj:
        lo, hi = EDGES[e]
        d = (CORNERS[hi] - CORNERS[lo]).astype(np.float64)
        out_dir[e] = d if inside[lo] else -d             # from the inside corner to the outside one
    tris, seen = [], set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        while True:
            a, b = adj[cur]
            nxt = a if a != prev else b
            if nxt == start:
                break
            loop.append(nxt)
            prev, cur = cur, nxt
        seen.update(loop)
        assert len(loop) >= 3, (case, loop)
        pts = np.array([mid[e] for e in loop])
        normal = np.zeros(3)
        for i in range(len(loop)):                       # Newell
            p, q = pts[i], pts[(i + 1) % len(loop)]
            normal += np.cross(p, q)
        if np.dot(normal, sum(out_dir[e] for e in loop)) < 0:
            loop = [loop[0]] + loop[:0:-1]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def build_tables():
    tri_lists = [_case_triangles(c) for c in range(256)]
    max_t = max(len(t) for t in tri_lists)
    ntri = np.array([len(t) for t in tri_lists], dtype=np.int32)
    tri = np.full((256, 3 * max_t), -1, dtype=np.int32)
    for c, t in enumerate(tri_lists):
        flat = [e for tr in t for e in tr]
        tri[c, : len(flat)] = flat
    return ntri, tri, max_t


MC_NTRI, MC_TRI, MC_MAX_TRI = build_tables()

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER_PATH = os.path.join(ROOT, "latent-diffusion-models-for-shape-sdfs_b200", "csrc", "mc_tables.h")
DATA_PATH = os.path.join(ROOT, "oracle", "mc_tables.npz")


def header_text() -> str:
    lines = ["// GENERATED by `python tools/gen_mc_tables.py` - do not edit.  Marching-cubes case tables built from the",
             "// face-segment construction described in tools/gen_mc_tables.py (corner i at (i&1, (i>>1)&1, i>>2); edge",
             "// e = 4*axis + k joins the k-th corner whose axis bit is 0 to that corner | (1 << axis)).",
             "#pragma once", "namespace sdfb {", f"constexpr int kMcMaxTri = {MC_MAX_TRI};",
             "// corner `lo` of each edge (hi = lo | (1 << (e >> 2)))",
             "__device__ __constant__ unsigned char kMcEdgeLo[12] = {" + ", ".join(str(int(a)) for a, _ in EDGES) + "};",
             "__device__ __constant__ unsigned char kMcNumTri[256] = {"]
    for r in range(0, 256, 32):
        lines.append("    " + ", ".join(str(int(v)) for v in MC_NTRI[r:r + 32]) + ",")
    lines.append("};")
    lines.append(f"// edge ids of the triangles' corners, {3 * MC_MAX_TRI} per case, 255 = unused")
    lines.append(f"__device__ __constant__ unsigned char kMcTri[256][{3 * MC_MAX_TRI}] = {{")
    for c in range(256):
        lines.append("    {" + ", ".join(str(int(v) if v >= 0 else 255) for v in MC_TRI[c]) + "},")
    lines.append("};")
    lines.append("}  // namespace sdfb")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    with open(HEADER_PATH, "w") as f:
        f.write(header_text())
    np.savez(DATA_PATH, ntri=MC_NTRI, tri=MC_TRI)
    print(f"wrote {HEADER_PATH}: max {MC_MAX_TRI} triangles per cell, {int(MC_NTRI.sum())} triangles over the 256 cases")
