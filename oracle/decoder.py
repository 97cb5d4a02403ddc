"""Oracle SDF auto-decoder: fp32 (or fp64) torch-CPU restatement and the
low-precision-emulating variant the tensor-core kernel is gated against.

Reference: none (`/root/reference/README.md:1`); follows SURVEY.md section 8(a)
rows A2/A3.  Test infrastructure only - never imported by the product path.

Network (DeepSDF-style):  in = concat(z[256], xyz[3]) = 259
    h0 = relu(L0 in)            259 -> 512
    h1 = relu(L1 h0)            512 -> 512
    h2 = relu(L2 h1)            512 -> 512
    h3 = relu(L3 h2)            512 -> 253
    h4 = relu(L4 concat(h3, in))  512 -> 512      (skip connection)
    h5..h7 = relu(L5..L7 .)     512 -> 512
    sdf = tanh(L8 h7)           512 -> 1
"""
from __future__ import annotations

import numpy as np
import torch

from .grid import grid_points
from .weights import DEC_LATENT, DEC_SKIP_OUT, decoder_weights


def _as_t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def decoder_forward(latent, xyz, params=None, dtype=torch.float32, chunk: int = 1 << 16):
    """Decoder(latent, xyz) -> sdf.  latent [256] (one shape) or [M,256]; xyz [M,3].

    Plain dense evaluation exactly as written in the module docstring, in
    ``dtype`` (float32 = the oracle; float64 = its self-check).
    Returns a float32/float64 numpy array [M].
    """
    params = decoder_weights() if params is None else params
    W = [_as_t(w, dtype) for w, _ in params]
    B = [_as_t(b, dtype) for _, b in params]
    xyz_t = _as_t(xyz, dtype).reshape(-1, 3)
    lat = _as_t(latent, dtype)
    M = xyz_t.shape[0]
    out = torch.empty(M, dtype=dtype)
    with torch.no_grad():
        for s in range(0, M, chunk):
            e = min(M, s + chunk)
            x = xyz_t[s:e]
            z = lat.expand(e - s, DEC_LATENT) if lat.ndim == 1 else lat[s:e]
            inp = torch.cat([z, x], dim=1)
            h = torch.relu(inp @ W[0].T + B[0])
            h = torch.relu(h @ W[1].T + B[1])
            h = torch.relu(h @ W[2].T + B[2])
            h = torch.relu(h @ W[3].T + B[3])
            h = torch.relu(torch.cat([h, inp], dim=1) @ W[4].T + B[4])
            h = torch.relu(h @ W[5].T + B[5])
            h = torch.relu(h @ W[6].T + B[6])
            h = torch.relu(h @ W[7].T + B[7])
            out[s:e] = torch.tanh(h @ W[8].T + B[8]).squeeze(1)
    return out.numpy()


def decode_grid(latent, res: int, z0: int = 0, z1: int | None = None, params=None,
                dtype=torch.float32):
    """decode_grid(z, res) -> sdf[z1-z0, res, res] (C-contiguous [z,y,x])."""
    z1 = res if z1 is None else z1
    sdf = decoder_forward(latent, grid_points(res, z0, z1), params=params, dtype=dtype)
    return sdf.reshape(z1 - z0, res, res)


def _round_to(t: torch.Tensor, lowp: torch.dtype) -> torch.Tensor:
    return t.to(lowp).to(torch.float32)


def decoder_forward_lowp(latent, xyz, params=None, lowp=torch.bfloat16, chunk: int = 1 << 16,
                         preacts: list | None = None):
    """The arithmetic the tensor-core kernel performs, emulated on the CPU.

    * per-shape constants are folded in fp32: bias0' = b0 + W0[:, :256] z and
      bias4' = b4 + W4[:, 253:509] z;
    * the xyz columns of L0 stay fp32 (the kernel evaluates the 3-wide first
      layer with fp32 FMAs): the geometry enters h0 unrounded;
    * L4 consumes [h3 | xyz] as ONE ``lowp`` operand: the coordinates ride in the
      three padding columns of h3's 256-wide slot, rounded to ``lowp`` like every
      other input of that layer, against W4's xyz columns rounded to ``lowp``;
    * every other weight is rounded to ``lowp``; activations h0..h6 are rounded
      to ``lowp`` after the ReLU; products accumulate in fp32;
    * h7 stays fp32 and the 512 -> 1 head and tanh are evaluated in fp32.
    Only a single shared latent ([256]) is supported, as in the kernel.

    ``preacts`` (diagnostics): if a list is passed, the fp32 pre-activations (before ReLU)
    of layers 1..7 of the LAST chunk are appended to it, [m, fout] each.
    """
    params = decoder_weights() if params is None else params
    f32 = torch.float32
    W = [_as_t(w, f32) for w, _ in params]
    B = [_as_t(b, f32) for _, b in params]
    z = _as_t(latent, f32).reshape(DEC_LATENT)
    xyz_t = _as_t(xyz, f32).reshape(-1, 3)
    L, S = DEC_LATENT, DEC_SKIP_OUT
    bias0 = B[0] + W[0][:, :L] @ z
    bias4 = B[4] + W[4][:, S:S + L] @ z
    W0x = W[0][:, L:L + 3]
    W4x = _round_to(W[4][:, S + L:S + L + 3], lowp)
    W4h = _round_to(W[4][:, :S], lowp)
    Wq = {i: _round_to(W[i], lowp) for i in (1, 2, 3, 5, 6, 7)}
    M = xyz_t.shape[0]
    out = torch.empty(M, dtype=f32)
    with torch.no_grad():
        for s in range(0, M, chunk):
            e = min(M, s + chunk)
            x = xyz_t[s:e]
            pre = []
            h = _round_to(torch.relu(x @ W0x.T + bias0), lowp)
            for li in (1, 2, 3):
                pre.append(h @ Wq[li].T + B[li])
                h = _round_to(torch.relu(pre[-1]), lowp)
            pre.append(h @ W4h.T + _round_to(x, lowp) @ W4x.T + bias4)
            h = _round_to(torch.relu(pre[-1]), lowp)
            for li in (5, 6):
                pre.append(h @ Wq[li].T + B[li])
                h = _round_to(torch.relu(pre[-1]), lowp)
            pre.append(h @ Wq[7].T + B[7])
            h = torch.relu(pre[-1])
            out[s:e] = torch.tanh(h @ W[8].T + B[8]).squeeze(1)
    if preacts is not None:
        preacts.extend(p.numpy() for p in pre)
    return out.numpy()


def decoder_vjp_latent(latent, xyz, dLdy, params=None, dtype=torch.float64):
    """grad[256] = sum_m dLdy[m] * d sdf(latent, xyz_m) / d latent, by torch autograd on the plain dense
    forward (the oracle of the CUDA backward path).  Returns (grad, sdf) numpy."""
    params = decoder_weights() if params is None else params
    W = [_as_t(w, dtype) for w, _ in params]
    B = [_as_t(b, dtype) for _, b in params]
    x = _as_t(xyz, dtype).reshape(-1, 3)
    up = _as_t(dLdy, dtype).reshape(-1)
    z = _as_t(latent, dtype).reshape(DEC_LATENT).clone().requires_grad_(True)
    inp = torch.cat([z.expand(x.shape[0], DEC_LATENT), x], dim=1)
    h = torch.relu(inp @ W[0].T + B[0])
    h = torch.relu(h @ W[1].T + B[1])
    h = torch.relu(h @ W[2].T + B[2])
    h = torch.relu(h @ W[3].T + B[3])
    h = torch.relu(torch.cat([h, inp], dim=1) @ W[4].T + B[4])
    h = torch.relu(h @ W[5].T + B[5])
    h = torch.relu(h @ W[6].T + B[6])
    h = torch.relu(h @ W[7].T + B[7])
    y = torch.tanh(h @ W[8].T + B[8]).squeeze(1)
    (y * up).sum().backward()
    return z.grad.detach().numpy(), y.detach().numpy()


def decoder_vjp_latent_lowp(latent, xyz, dLdy, params=None, lowp=torch.bfloat16):
    """The arithmetic of the tensor-core backward kernel (fused_decoder_kernel<., BWD>), emulated on the CPU:
    grad[256] = sum_m dLdy[m] d sdf_m / d latent with

    * the forward of ``decoder_forward_lowp`` (ReLU masks = pre-activation > 0);
    * g_m = (dLdy[m] 2^-e)(1 - y_m^2) in fp32, 2^e the power of two just above max |dLdy| (undone at the end:
      the deltas then sit in the same range whatever the caller's loss scale); the first backward operand is
      lowp(w8[n]) where h7's pre-activation is positive, and g_m multiplies the fp32 result of that product
      (it commutes with it), i.e. delta6 = lowp(g_m mask6 * ((mask7 * lowp(w8)) @ lowp(W7)));
    * delta_{l-1} = lowp(mask_{l-1} * (delta_l @ lowp(W_l))) with fp32 accumulation (the skip layer hands only
      its 253 hidden columns down);
    * the two column sums the latent sees, sum_m delta4 and sum_m delta0, are taken over the UNROUNDED fp32
      values, and contracted with the fp32 latent columns of W4 and W0.
    Returns (grad, sdf) as numpy arrays."""
    params = decoder_weights() if params is None else params
    f32 = torch.float32
    W = [_as_t(w, f32) for w, _ in params]
    B = [_as_t(b, f32) for _, b in params]
    z = _as_t(latent, f32).reshape(DEC_LATENT)
    x = _as_t(xyz, f32).reshape(-1, 3)
    up = _as_t(dLdy, f32).reshape(-1)
    L, S = DEC_LATENT, DEC_SKIP_OUT
    rnd = lambda t: _round_to(t, lowp)
    Wq = {i: rnd(W[i]) for i in (1, 2, 3, 5, 6, 7)}
    W4h, W4x = rnd(W[4][:, :S]), rnd(W[4][:, S + L:S + L + 3])
    with torch.no_grad():
        pre = [x @ W[0][:, L:L + 3].T + (B[0] + W[0][:, :L] @ z)]
        h = rnd(torch.relu(pre[0]))
        for li in (1, 2, 3):
            pre.append(h @ Wq[li].T + B[li])
            h = rnd(torch.relu(pre[-1]))
        pre.append(h @ W4h.T + rnd(x) @ W4x.T + (B[4] + W[4][:, S:S + L] @ z))
        h = rnd(torch.relu(pre[-1]))
        for li in (5, 6):
            pre.append(h @ Wq[li].T + B[li])
            h = rnd(torch.relu(pre[-1]))
        pre.append(h @ Wq[7].T + B[7])
        y = torch.tanh(torch.relu(pre[7]) @ W[8].T + B[8]).squeeze(1)
        mask = [(p > 0).to(f32) for p in pre]
        amax = float(up.abs().max()) if up.numel() else 0.0     # the kernel's power-of-two scaling of the upstream gradient
        ex = int(np.clip(np.frexp(np.float32(amax))[1], -100, 100)) if 0.0 < amax <= 3.0e38 else 0
        g = (up * float(2.0 ** -ex)) * (1.0 - y * y)
        d = rnd(W[8][0])[None, :] * mask[7]
        d = rnd(g[:, None] * ((d @ Wq[7]) * mask[6]))
        d = rnd((d @ Wq[6]) * mask[5])
        d4 = (d @ Wq[5]) * mask[4]
        d = rnd((rnd(d4) @ W4h) * mask[3])
        d = rnd((d @ Wq[3]) * mask[2])
        d = rnd((d @ Wq[2]) * mask[1])
        d0 = (d @ Wq[1]) * mask[0]
        grad = (d0.sum(0) @ W[0][:, :L] + d4.sum(0) @ W[4][:, S:S + L]) * float(2.0 ** ex)
    return grad.numpy(), y.numpy()


def fit_loss_grad_lowp(latent, xyz, sdf_target, clamp: float = 0.1, params=None, lowp=torch.bfloat16):
    """(loss, grad[256]) of the auto-decoder fitting loss mean_m |clamp(sdf_m) - clamp(target_m)| w.r.t. the latent, with
    the tensor-core kernel's arithmetic (decoder_forward_lowp / decoder_vjp_latent_lowp); the oracle of
    sdfb_decoder_fit_loss_grad.  sign(0) = 0 and the clamp's gradient is 1 strictly inside (-clamp, clamp)."""
    y = decoder_forward_lowp(latent, xyz, params=params, lowp=lowp)
    t = np.clip(np.asarray(sdf_target, dtype=np.float32).reshape(-1), -clamp, clamp)
    diff = np.clip(y, -clamp, clamp) - t
    M = max(y.shape[0], 1)
    up = (np.sign(diff) * ((y > -clamp) & (y < clamp))).astype(np.float32) / np.float32(M)
    grad, _ = decoder_vjp_latent_lowp(latent, xyz, up, params=params, lowp=lowp)
    return float(np.abs(diff).astype(np.float64).sum() / M), grad
