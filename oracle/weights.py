"""Frozen, seed-generated parameters of the oracle models (test infrastructure).

Reference: none exists (`/root/reference/README.md:1` is a title).  Shapes
follow SURVEY.md section 8(a): a DeepSDF-style auto-decoder
(259 -> 512 x3 -> 253 (+259 skip) -> 512 x4 -> 1) and an MLP denoiser
(512 -> 1024 x4 -> 256).

The parameters are produced by ``numpy.random.RandomState`` (the legacy
generator, whose stream numpy guarantees never to change) so that this
container and the GPU box build bit-identical weights without shipping
megabytes of fixtures.  ``weights_sha256`` pins them; the expected digests are
stored in ``tests/golden/``.

Decoder init is "geometric" (sphere-like field) rather than ``nn.Linear``'s
default, because the default gives an SDF with no zero crossing inside
[-1,1]^3 and every geometric test would be vacuous.
The head bias is a literal calibrated once so that about a quarter of the 64^3
grid lies inside the surface.
"""
from __future__ import annotations

import hashlib
from functools import lru_cache

import numpy as np

DEC_LATENT = 256
DEC_HIDDEN = 512
DEC_IN = DEC_LATENT + 3                 # concat(z, xyz)
DEC_SKIP_OUT = DEC_HIDDEN - DEC_IN      # 253: layer 3 emits this, so the skip concat is 512 wide
# (fan_in, fan_out) of the nine linear layers.
DEC_LAYER_DIMS = (
    (DEC_IN, DEC_HIDDEN),
    (DEC_HIDDEN, DEC_HIDDEN),
    (DEC_HIDDEN, DEC_HIDDEN),
    (DEC_HIDDEN, DEC_SKIP_OUT),
    (DEC_HIDDEN, DEC_HIDDEN),           # input = concat(h3[253], z[256], xyz[3])
    (DEC_HIDDEN, DEC_HIDDEN),
    (DEC_HIDDEN, DEC_HIDDEN),
    (DEC_HIDDEN, DEC_HIDDEN),
    (DEC_HIDDEN, 1),
)

DDPM_T = 1000
DDPM_LATENT = 256
DDPM_TEMB = 256
DDPM_HIDDEN = 1024
DDPM_LAYER_DIMS = (
    (DDPM_LATENT + DDPM_TEMB, DDPM_HIDDEN),
    (DDPM_HIDDEN, DDPM_HIDDEN),
    (DDPM_HIDDEN, DDPM_HIDDEN),
    (DDPM_HIDDEN, DDPM_HIDDEN),
    (DDPM_HIDDEN, DDPM_LATENT),
)

DEC_WEIGHT_SEED = 0
LATENT_SEED = 1
DDPM_WEIGHT_SEED = 3

# Calibrated once with oracle/make_golden.py --calibrate (fp32 oracle, 64^3,
# default latent): puts ~25% of the grid nodes inside the zero level set.
DEC_HEAD_BIAS = -2.7683735


@lru_cache(maxsize=None)
def decoder_weights():
    """List of (W[out,in] float32, b[out] float32) for the nine decoder layers."""
    rs = np.random.RandomState(DEC_WEIGHT_SEED)
    params = []
    for li, (fin, fout) in enumerate(DEC_LAYER_DIMS):
        if li < 8:
            std = np.sqrt(2.0) / np.sqrt(fout)
            w = rs.standard_normal((fout, fin)) * std
            b = rs.standard_normal(fout) * 0.05
            if li == 0:
                # xyz columns carry the geometry; the latent columns perturb it.
                w[:, :DEC_LATENT] *= 0.5
            if li == 4:
                # skip layer: damp the re-injected input so depth still matters
                w[:, DEC_SKIP_OUT:] *= 0.5
        else:
            mean = np.sqrt(np.pi) / np.sqrt(fin)
            w = mean + rs.standard_normal((fout, fin)) * (0.6 * mean)
            b = np.full(fout, DEC_HEAD_BIAS)
        params.append((np.ascontiguousarray(w, dtype=np.float32),
                       np.ascontiguousarray(b, dtype=np.float32)))
    return params


@lru_cache(maxsize=None)
def ddpm_weights():
    """List of (W[out,in] float32, b[out] float32) for the five denoiser layers."""
    rs = np.random.RandomState(DDPM_WEIGHT_SEED)
    params = []
    for fin, fout in DDPM_LAYER_DIMS:
        std = 1.0 / np.sqrt(fin)
        w = rs.standard_normal((fout, fin)) * std
        b = rs.standard_normal(fout) * 0.02
        params.append((np.ascontiguousarray(w, dtype=np.float32),
                       np.ascontiguousarray(b, dtype=np.float32)))
    return params


def default_latent(index: int = 0) -> np.ndarray:
    """Latent code z ~ N(0, 1/256) (DeepSDF's code-init scale), float32 [256]."""
    rs = np.random.RandomState(LATENT_SEED + 1000 * index)
    return (rs.standard_normal(DEC_LATENT) / 16.0).astype(np.float32)


def flatten_params(params) -> np.ndarray:
    """Flat float32 blob in the C-ABI order: W0, b0, W1, b1, ... (row-major W[out,in])."""
    return np.concatenate([np.concatenate([w.ravel(), b.ravel()]) for w, b in params]).astype(np.float32)


def weights_sha256(params) -> str:
    return hashlib.sha256(flatten_params(params).tobytes()).hexdigest()
