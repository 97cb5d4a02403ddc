"""Timing of the training steps for bench.py: a DDPM denoiser step and an auto-decoder step
on the tensor pipe, with the FLOPs the products actually execute (padded shapes included)."""
from __future__ import annotations

import statistics

import torch

from . import synthetic
from .api import DDPMTrainer, DecoderTrainer

# products of one DDPM step per batch row: forward (512->1024, 3 x 1024->1024, 1024->256), backward-data (the same without the
# first layer), weight gradients (the same as forward)
_DDPM_FWD = 2 * (512 * 1024 + 3 * 1024 * 1024 + 1024 * 256)
DDPM_TRAIN_FLOP_PER_ROW = _DDPM_FWD + (_DDPM_FWD - 2 * 512 * 1024) + _DDPM_FWD
# decoder step per point, as executed (inputs padded to 320, layer 3 to 256 outputs): forward / weight-gradient products
_DEC_FWD = 2 * (320 * 512 + 2 * 512 * 512 + 512 * 256 + 4 * 512 * 512)
_DEC_BWD = 2 * (3 * 512 * 512 + 512 * 256 + 256 * 512 + 2 * 512 * 512)      # delta_6..delta_0
DEC_TRAIN_FLOP_PER_POINT = _DEC_FWD + _DEC_BWD + _DEC_FWD


def bench_training_legs(dec, dev, rank, world, precision, max_over_ranks, barrier, peaks):
    out = {}
    g = torch.Generator(device=dev).manual_seed(900 + rank)
    # ---- DDPM training step, 4096 latents per step
    n = 4096
    tr = DDPMTrainer(synthetic.ddpm_params(), device=dev, precision=precision)
    x0 = torch.randn((n, 256), generator=g, device=dev).clamp_(-1, 1)
    eps = torch.randn((n, 256), generator=g, device=dev)
    t = torch.randint(0, 1000, (n,), generator=g, device=dev, dtype=torch.int32)
    first = float(tr.step(x0, t, eps, lr=1e-4).item())
    for _ in range(3):
        tr.step(x0, t, eps, lr=1e-4)
    barrier()
    ms = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            loss = tr.step(x0, t, eps, lr=1e-4)
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b) / 10)
    m = max_over_ranks(statistics.median(ms))
    out["ddpm_train_step"] = {"workload": f"per GPU: DDPMTrainer.step on {n} latents (noising, 5-layer denoiser forward, MSE, backward, weight "
                                          "gradients, Adam on 3.9 M parameters; ~30 launches)",
                              "ms_per_step": m, "latents_per_s": world * n / (m * 1e-3),
                              "achieved_tflops_per_gpu": n * DDPM_TRAIN_FLOP_PER_ROW / (m * 1e-3) / 1e12,
                              "loss_first": first, "loss_after_54_steps_same_batch": float(loss.item())}
    tr.close()
    # ---- auto-decoder training step: 64 shapes x 8192 samples
    B, P = 64, 8192
    dt = DecoderTrainer(synthetic.decoder_params(), device=dev, precision=precision)
    lat = torch.stack([torch.from_numpy(synthetic.latent(i)) for i in range(B)]).to(dev)
    xyz = torch.rand((B, P, 3), generator=g, device=dev) * 2 - 1
    tgt = torch.stack([dec(torch.from_numpy(synthetic.latent(100 + i)).to(dev), xyz[i]) for i in range(B)])
    first = float(dt.step(lat, xyz, tgt, lr=1e-4).item())
    dt.step(lat, xyz, tgt, lr=1e-4)
    barrier()
    ms = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = dt.step(lat, xyz, tgt, lr=1e-4)
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    m = max_over_ranks(statistics.median(ms))
    out["decoder_train_step"] = {"workload": f"per GPU: DecoderTrainer.step on {B} shapes x {P} samples (forward, clamped-L1, backward, weight "
                                             "gradients of all nine layers, Adam on 1.8 M parameters; layer by layer, activations and deltas in 16 bits)",
                                 "ms_per_step": m, "points_per_s": world * B * P / (m * 1e-3),
                                 "achieved_tflops_per_gpu": B * P * DEC_TRAIN_FLOP_PER_POINT / (m * 1e-3) / 1e12,
                                 "loss_first": first, "loss_after_7_steps_same_batch": float(loss.item())}
    for k in out:
        out[k]["frac_of_burst_peak"] = out[k]["achieved_tflops_per_gpu"] / peaks["burst"]
    dt.close()
    return out
