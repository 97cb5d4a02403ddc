"""Oracle of the training steps (second half): DDPM denoiser training step and decoder
weight gradients.

Reference: none (`/root/reference/README.md:1` is a title); the loss is the standard epsilon-prediction objective of the
latent DDPM frozen in oracle/ddpm.py (x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) eps, loss = mean (eps_hat - eps)^2) and
Adam with bias correction.  Two evaluations of the gradient:
  * `ddpm_train_grads`      : torch autograd, fp64 by default - the mathematical answer;
  * `ddpm_train_grads_lowp` : what the tensor-core step computes - weights, activations and deltas rounded to the 16-bit
                              operand type at every product, fp32 accumulation, ReLU masks from the rounded activations.
Test infrastructure only.
"""
from __future__ import annotations

import numpy as np
import torch

from .ddpm import ddpm_schedule, time_embedding  # noqa: F401
from .weights import DDPM_T, ddpm_weights


def _abar_coefs():
    beta = np.linspace(1e-4, 0.02, DDPM_T, dtype=np.float64)
    abar = np.cumprod(1.0 - beta)
    return np.sqrt(abar).astype(np.float32), np.sqrt(1.0 - abar).astype(np.float32)


def ddpm_noised_input(x0, t, eps):
    """[x_t | temb(t)] float32 [n,512] with x_t = fl(fl(a_t x0) + fl(b_t eps)) as the kernel computes it."""
    a, b = _abar_coefs()
    x0 = np.asarray(x0, np.float32)
    eps = np.asarray(eps, np.float32)
    t = np.asarray(t, np.int64)
    xt = (a[t][:, None] * x0).astype(np.float32) + (b[t][:, None] * eps).astype(np.float32)
    return np.concatenate([xt.astype(np.float32), time_embedding(t)], axis=1).astype(np.float32)


def ddpm_train_grads(x0, t, eps, params=None, dtype=torch.float64):
    """(loss, [(dW_l, db_l)]) by autograd."""
    params = ddpm_weights() if params is None else params
    W = [torch.tensor(np.asarray(w), dtype=dtype, requires_grad=True) for w, _ in params]
    B = [torch.tensor(np.asarray(b), dtype=dtype, requires_grad=True) for _, b in params]
    h = torch.tensor(ddpm_noised_input(x0, t, eps), dtype=dtype)
    for i in range(4):
        h = torch.relu(h @ W[i].T + B[i])
    out = h @ W[4].T + B[4]
    loss = ((out - torch.tensor(np.asarray(eps), dtype=dtype)) ** 2).mean()
    loss.backward()
    return float(loss.detach()), [(w.grad.numpy(), b.grad.numpy()) for w, b in zip(W, B)]


def ddpm_train_grads_lowp(x0, t, eps, params=None, lowp=torch.bfloat16):
    """(loss, [(dW_l, db_l)]) with the tensor-core step's roundings (see the module docstring)."""
    params = ddpm_weights() if params is None else params
    f32 = torch.float32

    def rnd(x):
        return x.to(lowp).to(f32)

    W = [torch.tensor(np.asarray(w), dtype=f32) for w, _ in params]
    B = [torch.tensor(np.asarray(b), dtype=f32) for _, b in params]
    Wl = [rnd(w) for w in W]
    with torch.no_grad():
        acts = [rnd(torch.tensor(ddpm_noised_input(x0, t, eps)))]
        for i in range(4):
            acts.append(rnd(torch.relu(acts[i] @ Wl[i].T + B[i])))
        out = acts[4] @ Wl[4].T + B[4]
        d = out - torch.tensor(np.asarray(eps), dtype=f32)
        n = d.shape[0]
        loss = float((d.double() ** 2).mean())
        scale = 2.0 / (n * 256)
        delta = rnd(d)
        grads = [None] * 5
        for l in range(4, -1, -1):
            prev = rnd((delta @ Wl[l]) * (acts[l] > 0)) if l > 0 else None
            grads[l] = ((delta.T @ acts[l]).numpy() * np.float32(scale), delta.sum(dim=0).numpy() * np.float32(scale))
            delta = prev
    return loss, grads


def adam_step(params, grads, m, v, step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """One Adam update with bias correction on flat float32 arrays (in the kernel's operation order)."""
    g = np.asarray(grads, np.float32)
    m = (np.float32(beta1) * m + np.float32(1.0 - beta1) * g).astype(np.float32)
    v = (np.float32(beta2) * v + np.float32(1.0 - beta2) * g * g).astype(np.float32)
    c1 = np.float32(1.0 - np.float32(beta1) ** np.float32(step))
    c2 = np.float32(1.0 - np.float32(beta2) ** np.float32(step))
    p = (params - np.float32(lr) * (m / c1) / (np.sqrt(v / c2) + np.float32(eps))).astype(np.float32)
    return p, m, v


def flatten_grads(grads) -> np.ndarray:
    """[(dW, db)] -> the parameter blob's layout (W0, b0, W1, b1, ...)."""
    return np.concatenate([np.concatenate([np.asarray(w, np.float32).ravel(), np.asarray(b, np.float32).ravel()]) for w, b in grads])


# ---- auto-decoder training step: gradients of the decoder's weights ------------------------------------------------------
def _decoder_inputs(latents, xyz):
    lat = np.asarray(latents, np.float32)
    pts = np.asarray(xyz, np.float32)
    B, P = pts.shape[0], pts.shape[1]
    return np.concatenate([np.repeat(lat[:, None, :], P, axis=1), pts], axis=2).reshape(B * P, 259)


def decoder_train_grads(latents, xyz, target, clamp: float = 0.1, params=None, dtype=torch.float64):
    """(loss, [(dW_l, db_l)] for the nine layers, sdf [B*P]) by autograd: loss = mean |clamp(sdf) - clamp(target)| over a
    batch of shapes (latents [B,256], xyz [B,P,3], target [B,P]); the network is oracle/decoder.py's."""
    from .weights import decoder_weights
    params = decoder_weights() if params is None else params
    W = [torch.tensor(np.asarray(w), dtype=dtype, requires_grad=True) for w, _ in params]
    Bs = [torch.tensor(np.asarray(b), dtype=dtype, requires_grad=True) for _, b in params]
    x = torch.tensor(_decoder_inputs(latents, xyz), dtype=dtype)
    h = x
    for l in range(8):
        if l == 4:
            h = torch.cat([h, x], dim=1)
        h = torch.relu(h @ W[l].T + Bs[l])
    y = torch.tanh(h @ W[8].T + Bs[8])[:, 0]
    tgt = torch.tensor(np.asarray(target, np.float32).reshape(-1), dtype=dtype)
    loss = (torch.clamp(y, -clamp, clamp) - torch.clamp(tgt, -clamp, clamp)).abs().mean()
    loss.backward()
    return float(loss.detach()), [(w.grad.numpy(), b.grad.numpy()) for w, b in zip(W, Bs)], y.detach().numpy()


def decoder_train_grads_lowp(latents, xyz, target, clamp: float = 0.1, params=None, lowp=torch.bfloat16):
    """The same with the tensor-core step's roundings: inputs, weights (layers 0-7), activations and deltas rounded to the 16-bit
    operand type at every product, fp32 accumulation; the head (layer 8) in fp32 on the rounded h7; ReLU masks from the rounded
    activations; the upstream gradient sign(clamp(y) - clamp(t)) [|y| < clamp] / M formed from the value just computed."""
    from .weights import decoder_weights
    params = decoder_weights() if params is None else params
    f32 = torch.float32

    def rnd(v):
        return v.to(lowp).to(f32)

    W = [torch.tensor(np.asarray(w), dtype=f32) for w, _ in params]
    Bs = [torch.tensor(np.asarray(b), dtype=f32) for _, b in params]
    Wl = [rnd(w) for w in W[:8]]
    with torch.no_grad():
        x = rnd(torch.tensor(_decoder_inputs(latents, xyz)))
        acts = [x]                                   # acts[l] = input of layer l
        h = x
        for l in range(8):
            if l == 4:
                h = torch.cat([h, x], dim=1)
                acts[4] = h
            h = rnd(torch.relu(h @ Wl[l].T + Bs[l]))
            acts.append(h)
        y = torch.tanh(acts[8] @ W[8].T + Bs[8])[:, 0]
        tgt = torch.tensor(np.asarray(target, np.float32).reshape(-1))
        M = y.shape[0]
        diff = torch.clamp(y, -clamp, clamp) - torch.clamp(tgt, -clamp, clamp)
        loss = float(diff.abs().double().mean())
        up = torch.sign(diff) * ((y > -clamp) & (y < clamp))
        d8 = up * (1.0 - y * y)
        inv_m = np.float32(1.0 / M)
        grads = [None] * 9
        grads[8] = (((d8[:, None] * acts[8]).sum(dim=0)[None, :]).numpy() * inv_m, np.array([float(d8.sum())], np.float32) * inv_m)
        delta = rnd((d8[:, None] * W[8]) * (acts[8] > 0))                    # delta_7
        for l in range(7, -1, -1):
            grads[l] = ((delta.T @ acts[l]).numpy() * inv_m, delta.sum(dim=0).numpy() * inv_m)
            if l > 0:
                back = delta @ Wl[l]                                        # gradient w.r.t. layer l's input
                if l == 4:
                    back = back[:, :253]
                delta = rnd(back * (acts[l][:, :back.shape[1]] > 0))
    return loss, grads, y.numpy()
