"""Seeded random-init parameters of the two networks, for benchmarks and demos.

There are no checkpoints to load (no network, and the upstream repository has no code:
`/root/reference/README.md:1`), so `bench.py` runs on random-init weights of the architecture,
as its contract says.  The recipe (numpy's legacy ``RandomState`` stream, "geometric" decoder
init so that the field has a zero level set) is the one the test oracle freezes in
``oracle/weights.py``; ``tests/test_oracle.py`` asserts that the two produce identical bytes, so
nothing on the product or benchmark path needs to import ``oracle``.
"""
from __future__ import annotations

import numpy as np

LATENT = 256
DEC_DIMS = ((259, 512), (512, 512), (512, 512), (512, 253), (512, 512), (512, 512), (512, 512), (512, 512), (512, 1))
DDPM_DIMS = ((512, 1024), (1024, 1024), (1024, 1024), (1024, 1024), (1024, 256))
DEC_HEAD_BIAS = -2.7683735      # puts ~25 % of the 64^3 nodes of latent 0 inside the surface


def decoder_params(seed: int = 0) -> np.ndarray:
    """Flat float32 blob W0,b0,...,W8,b8 (row-major W[out][in]) - the C ABI's input format."""
    rs = np.random.RandomState(seed)
    parts = []
    for li, (fin, fout) in enumerate(DEC_DIMS):
        if li < 8:
            w = rs.standard_normal((fout, fin)) * (np.sqrt(2.0) / np.sqrt(fout))
            b = rs.standard_normal(fout) * 0.05
            if li == 0:
                w[:, :LATENT] *= 0.5
            if li == 4:
                w[:, 253:] *= 0.5
        else:
            mean = np.sqrt(np.pi) / np.sqrt(fin)
            w = mean + rs.standard_normal((fout, fin)) * (0.6 * mean)
            b = np.full(fout, DEC_HEAD_BIAS)
        parts += [np.asarray(w, dtype=np.float32).ravel(), np.asarray(b, dtype=np.float32).ravel()]
    return np.concatenate(parts).astype(np.float32)


def ddpm_params(seed: int = 3) -> np.ndarray:
    """Flat float32 blob W0,b0,...,W4,b4 of the denoiser."""
    rs = np.random.RandomState(seed)
    parts = []
    for fin, fout in DDPM_DIMS:
        w = rs.standard_normal((fout, fin)) * (1.0 / np.sqrt(fin))
        b = rs.standard_normal(fout) * 0.02
        parts += [np.asarray(w, dtype=np.float32).ravel(), np.asarray(b, dtype=np.float32).ravel()]
    return np.concatenate(parts).astype(np.float32)


def latent(index: int = 0) -> np.ndarray:
    """z ~ N(0, 1/256), float32 [256]."""
    rs = np.random.RandomState(1 + 1000 * index)
    return (rs.standard_normal(LATENT) / 16.0).astype(np.float32)
