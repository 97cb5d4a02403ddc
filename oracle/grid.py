"""Grid coordinates and the sign-change (narrow-band) mask - oracle definitions.

Reference: none (`/root/reference/README.md:1`); follows SURVEY.md section 8(a)
rows A1 and A4.  Test infrastructure only.
"""
from __future__ import annotations

import numpy as np


def axis_coords(res: int) -> np.ndarray:
    """A1: c_i = float32(2i - (res-1)) / float32(res-1), one correctly rounded divide.

    The numerator and denominator are small integers (exact in fp32), so a
    single IEEE divide gives the same bits on every conforming CPU and GPU.
    """
    if res < 2:
        raise ValueError("res must be >= 2")
    i = np.arange(res, dtype=np.int64)
    num = (2 * i - (res - 1)).astype(np.float32)
    den = np.float32(res - 1)
    return (num / den).astype(np.float32)


def grid_points(res: int, z0: int = 0, z1: int | None = None) -> np.ndarray:
    """xyz of planes [z0, z1) of the res^3 grid, float32 [(z1-z0)*res*res, 3].

    Query q = (iz*res + iy)*res + ix (x fastest); column order is (x, y, z).
    """
    z1 = res if z1 is None else z1
    c = axis_coords(res)
    zz, yy, xx = np.meshgrid(c[z0:z1], c, c, indexing="ij")
    return np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1).astype(np.float32)


def sign_change_mask(sdf: np.ndarray) -> np.ndarray:
    """A4: uint8 [(nz-1),(ny-1),(nx-1)], 1 where the cell's 8 corners are not all
    on one side.  inside(v) := v < 0 (so -0.0 and NaN count as outside)."""
    s = np.asarray(sdf)
    if s.ndim != 3:
        raise ValueError("sdf must be [nz, ny, nx]")
    inside = s < 0
    any_in = np.zeros(tuple(d - 1 for d in s.shape), dtype=bool)
    all_in = np.ones_like(any_in)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                c = inside[dz:s.shape[0] - 1 + dz, dy:s.shape[1] - 1 + dy, dx:s.shape[2] - 1 + dx]
                any_in |= c
                all_in &= c
    return (any_in & ~all_in).astype(np.uint8)
