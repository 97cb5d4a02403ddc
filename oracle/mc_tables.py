"""Marching-cubes case tables of the CPU oracle, as DATA (test infrastructure only).

The numbers live in oracle/mc_tables.npz, written by the product's build tool tools/gen_mc_tables.py together with the
header the GPU kernels include (csrc/mc_tables.h); tests/test_marching.py checks that both files match the generator and
checks the tables themselves against properties the construction does not know about.  Conventions:
corner i sits at (x, y, z) = (i & 1, (i >> 1) & 1, i >> 2); case index = sum(inside(corner i) << i), inside(v) := v < 0;
edge e = 4 * axis + k joins corner `lo` (the k-th corner, ascending, whose `axis` bit is 0) to lo | (1 << axis); vertices
are always interpolated from lo to hi.  No upstream source exists (/root/reference/README.md:1).
"""
from __future__ import annotations

import os

import numpy as np

CORNERS = np.array([[i & 1, (i >> 1) & 1, i >> 2] for i in range(8)], dtype=np.int64)
EDGES = np.array([(c, c | (1 << axis)) for axis in range(3) for c in range(8) if not (c >> axis) & 1], dtype=np.int64)   # [12, 2] (lo, hi)
EDGE_AXIS = np.repeat(np.arange(3), 4)

_data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mc_tables.npz"))
MC_NTRI = _data["ntri"].astype(np.int32)      # [256] triangles per case
MC_TRI = _data["tri"].astype(np.int32)        # [256, 15] edge ids, -1 = unused
MC_MAX_TRI = MC_TRI.shape[1] // 3
