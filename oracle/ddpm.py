"""Oracle latent-space DDPM sampler (epsilon-prediction, linear beta schedule,
1000 steps, x0-clipped posterior-mean update) and its low-precision variant.

Reference: none (`/root/reference/README.md:1`); follows SURVEY.md section 8(a)
rows A5-A7 (the x0-clipped form is the one frozen there, section 7 H4, because
the plain ancestral update amplifies rounding 157x and makes the 1e-4 parity
criterion meaningless).  Test infrastructure only.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
import torch

from .weights import DDPM_T, DDPM_LATENT, DDPM_TEMB, ddpm_weights

# The DDPM works on latents normalised to unit scale (x0 is clipped to [-1, 1]); a decoder latent is a sample times this
# factor - the N(0, 1/16^2) scale of the auto-decoder's codes (oracle/weights.py default_latent).  Decoding an unscaled
# sample saturates the field (tanh -> +1 everywhere, no surface), which would make config 4's geometric checks vacuous.
DDPM_LATENT_SCALE = 1.0 / 16.0


def to_decoder_latent(x0):
    """x0 [n,256] (DDPM space) -> decoder latents [n,256]."""
    return (np.asarray(x0, dtype=np.float32) * np.float32(DDPM_LATENT_SCALE)).astype(np.float32)


@lru_cache(maxsize=None)
def ddpm_schedule(T: int = DDPM_T):
    """Per-step coefficients, computed in fp64 and cast to fp32.

    Returns dict of float32 arrays [T]:
      sra  = 1/sqrt(abar_t)            srm1 = sqrt(1/abar_t - 1)
      c1   = beta_t sqrt(abar_{t-1}) / (1 - abar_t)
      c2   = (1 - abar_{t-1}) sqrt(alpha_t) / (1 - abar_t)
      sigma = sqrt(beta_t (1 - abar_{t-1}) / (1 - abar_t))      (0 at t = 0)
    """
    beta = np.linspace(1e-4, 0.02, T, dtype=np.float64)
    alpha = 1.0 - beta
    abar = np.cumprod(alpha)
    abar_prev = np.concatenate([[1.0], abar[:-1]])
    sched = {
        "sra": 1.0 / np.sqrt(abar),
        "srm1": np.sqrt(1.0 / abar - 1.0),
        "c1": beta * np.sqrt(abar_prev) / (1.0 - abar),
        "c2": (1.0 - abar_prev) * np.sqrt(alpha) / (1.0 - abar),
        "sigma": np.sqrt(beta * (1.0 - abar_prev) / (1.0 - abar)),
    }
    sched["sigma"][0] = 0.0
    return {k: v.astype(np.float32) for k, v in sched.items()}


def time_embedding(t, dim: int = DDPM_TEMB) -> np.ndarray:
    """Sinusoidal embedding [len(t), dim] = concat(sin(t f_i), cos(t f_i)),
    f_i = 10000^(-i/half), evaluated in fp64 and cast to fp32."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    half = dim // 2
    freqs = np.exp(-math.log(10000.0) * np.arange(half, dtype=np.float64) / half)
    ang = t[:, None] * freqs[None, :]
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1).astype(np.float32)


def _as_t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def denoiser_forward(x, t: int, params=None, dtype=torch.float32):
    """eps_hat(x_t, t): in = concat(x[256], temb(t)[256]) -> 1024 x4 (ReLU) -> 256."""
    params = ddpm_weights() if params is None else params
    W = [_as_t(w, dtype) for w, _ in params]
    B = [_as_t(b, dtype) for _, b in params]
    xt = _as_t(x, dtype).reshape(-1, DDPM_LATENT)
    te = _as_t(time_embedding(t), dtype).expand(xt.shape[0], DDPM_TEMB)
    with torch.no_grad():
        h = torch.cat([xt, te], dim=1)
        for i in range(4):
            h = torch.relu(h @ W[i].T + B[i])
        return h @ W[4].T + B[4]


def _round_to(t, lowp):
    return t.to(lowp).to(torch.float32)


def denoiser_forward_lowp(x, t: int, params=None, lowp=torch.bfloat16):
    """What the tensor-core denoiser computes: the time-embedding half of layer 0
    is folded into a per-step fp32 bias; x_t enters as an exact 2-way ``lowp``
    split (hi + lo); weights and hidden activations are rounded to ``lowp``;
    accumulation, the final layer's output and the update are fp32."""
    params = ddpm_weights() if params is None else params
    f32 = torch.float32
    W = [_as_t(w, f32) for w, _ in params]
    B = [_as_t(b, f32) for _, b in params]
    xt = _as_t(x, f32).reshape(-1, DDPM_LATENT)
    te = _as_t(time_embedding(t), f32)[0]
    with torch.no_grad():
        bias0 = B[0] + W[0][:, DDPM_LATENT:] @ te
        x_hi = _round_to(xt, lowp)
        x_lo = _round_to(xt - x_hi, lowp)
        W0x = _round_to(W[0][:, :DDPM_LATENT], lowp)
        h = _round_to(torch.relu(x_hi @ W0x.T + x_lo @ W0x.T + bias0), lowp)
        for i in (1, 2, 3):
            h = _round_to(torch.relu(h @ _round_to(W[i], lowp).T + B[i]), lowp)
        return h @ _round_to(W[4], lowp).T + B[4]


def ddpm_step(x, eps, noise, t: int, sched=None):
    """A7: x0 = clamp(sra x - srm1 eps, -1, 1);  x <- c1 x0 + c2 x + sigma noise."""
    sched = ddpm_schedule() if sched is None else sched
    dt = x.dtype
    c = {k: torch.tensor(float(v[t]), dtype=dt) for k, v in sched.items()}
    x0 = torch.clamp(c["sra"] * x - c["srm1"] * eps, -1.0, 1.0)
    out = c["c1"] * x0 + c["c2"] * x
    if t > 0:
        out = out + c["sigma"] * noise
    return out


def sample_latents(n: int, x_T, noise, params=None, dtype=torch.float32, lowp=None,
                   steps: int = DDPM_T):
    """sample_latents(n): loop t = steps-1 .. 0 from x_T [n,256] with the explicit
    noise stream ``noise`` [steps, n, 256] (noise[t] is consumed at step t; noise[0]
    is ignored).  ``lowp`` selects the tensor-core-emulating denoiser."""
    x = _as_t(x_T, dtype).reshape(n, DDPM_LATENT).clone()
    nz = _as_t(noise, dtype)
    sched = ddpm_schedule()
    for t in range(steps - 1, -1, -1):
        if lowp is None:
            eps = denoiser_forward(x, t, params=params, dtype=dtype)
        else:
            eps = denoiser_forward_lowp(x, t, params=params, lowp=lowp).to(dtype)
        x = ddpm_step(x, eps, nz[t], t, sched)
    return x.numpy()
