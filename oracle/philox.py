"""Oracle of the in-kernel noise generator: Philox4x32-10 counters -> standard normals.

Reference: none in the mounted tree (`/root/reference/README.md:1`); Philox4x32-10 is the
counter-based generator of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
restated here from the published algorithm and pinned by the Random123 known-answer vectors in
tests/test_oracle.py.  The product kernels (csrc/philox.cuh) use the same counter layout so that
the noise stream of a seeded sampling run can be reproduced on the CPU ("identical noise stream").
Test infrastructure only.

Counter layout (one Philox call = 4 consecutive columns of one latent at one step):
    ctr = (column // 4, latent index, t, 0x53444642)     key = (seed & 0xffffffff, seed >> 32)
    x_T uses t = steps (one past the largest noise index); noise[t] is consumed at step t.
Normals (Box-Muller, 24-bit uniforms so every intermediate is exact in float32 up to the log / sqrt /
sin / cos themselves):
    u1 = ((r_a >> 8) + 1) 2^-24 in (0, 1],  u2 = (r_b >> 8) 2^-24 in [0, 1)
    z_a = sqrt(-2 ln u1) cos(2 pi u2),  z_b = sqrt(-2 ln u1) sin(2 pi u2);   (r0, r1) and (r2, r3) are the two pairs.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
TAG = 0x53444642
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 array [..., 4]; key: (k0, k1) python ints.  Returns uint32 [..., 4]."""
    c = np.asarray(ctr, dtype=np.uint32)
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _box_muller(ra, rb):
    u1 = ((ra >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (rb >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = 2.0 * np.pi * u2.astype(np.float64)
    return (rad * np.cos(ang).astype(np.float32)).astype(np.float32), (rad * np.sin(ang).astype(np.float32)).astype(np.float32)


def philox_normal_rows(seed: int, n: int, t0: int, t1: int, width: int = 256, first_latent: int = 0) -> np.ndarray:
    """Normals [t1 - t0, n, width] float32 for steps t in [t0, t1) of latents [first_latent, first_latent + n)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    t = np.arange(t0, t1, dtype=np.uint32)[:, None, None]
    i = (np.arange(n, dtype=np.uint64) + np.uint64(first_latent)).astype(np.uint32)[None, :, None]
    g = np.arange(width // 4, dtype=np.uint32)[None, None, :]
    shape = (t1 - t0, n, width // 4)
    ctr = np.stack([np.broadcast_to(g, shape), np.broadcast_to(i, shape), np.broadcast_to(t, shape),
                    np.full(shape, TAG, dtype=np.uint32)], axis=-1)
    r = philox4x32_10(ctr, key)
    z0, z1 = _box_muller(r[..., 0], r[..., 1])
    z2, z3 = _box_muller(r[..., 2], r[..., 3])
    return np.stack([z0, z1, z2, z3], axis=-1).reshape(t1 - t0, n, width).astype(np.float32)


def philox_sampler_inputs(seed: int, n: int, steps: int):
    """(x_T [n,256], noise [steps,n,256]) of a seeded sampling run: x_T is row t = steps."""
    noise = philox_normal_rows(seed, n, 0, steps)
    x_T = philox_normal_rows(seed, n, steps, steps + 1)[0]
    return x_T, noise
