#!/usr/bin/env python
"""Device-timed numbers for every BASELINE.json config that fits one GPU, plus points mode and fp16
(CUDA events, 1 warm-up + median of 3; inputs resident in HBM).  Writes a small table to stdout."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package

pkg = load_package()
dev = torch.device("cuda:0")
dec = pkg.Decoder(pkg.synthetic.decoder_params(), device=dev)
ddpm = pkg.LatentDDPM(pkg.synthetic.ddpm_params(), device=dev, precision="bf16")
z = torch.from_numpy(pkg.synthetic.latent(0)).to(dev)
FLOP_Q = 2 * 6 * 512 * 512


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return statistics.median(ms)


rows = []
for res in (64, 128, 256, 512):
    out = torch.empty((res, res, res), device=dev)
    ms = timed(lambda: dec.decode_grid(z, res, out=out))
    rows.append((f"decode_grid {res}^3 bf16", ms, res ** 3 / ms / 1e3, res ** 3 * FLOP_Q / ms / 1e9))
    del out
out = torch.empty((256, 256, 256), device=dev)
ms = timed(lambda: dec.decode_grid(z, 256, out=out, precision="fp16"))
rows.append(("decode_grid 256^3 fp16", ms, 256 ** 3 / ms / 1e3, 256 ** 3 * FLOP_Q / ms / 1e9))
ms = timed(lambda: dec.decode_grid(z, 256, mask=True))
rows.append(("decode_grid 256^3 bf16 + mask", ms, 256 ** 3 / ms / 1e3, 256 ** 3 * FLOP_Q / ms / 1e9))
pts = torch.rand((256 ** 3, 3), device=dev) * 2 - 1
ms = timed(lambda: dec(z, pts))
rows.append(("Decoder(latent, xyz) 16.7M random points bf16", ms, 256 ** 3 / ms / 1e3, 256 ** 3 * FLOP_Q / ms / 1e9))
del pts, out
lat = torch.stack([torch.from_numpy(pkg.synthetic.latent(i)) for i in range(64)]).to(dev)
buf = torch.empty((64, 128, 128, 128), device=dev)
ms = timed(lambda: dec.decode_grid_batch(lat, 128, out=buf), reps=2)
q = 64 * 128 ** 3
rows.append(("config 3 on ONE GPU: 64 latents x 128^3 bf16", ms, q / ms / 1e3, q * FLOP_Q / ms / 1e9))
sdf, signs, _ = dec.decode_grid_bits(z, 256, mask=False)
ms = timed(lambda: pkg.extract_surface(sdf, 256, 0, sign_words=signs))
rows.append(("marching cubes 256^3 (count + scan + generate)", ms, 255 ** 3 / ms / 1e3, 0.0))
# latent gradient (row N4): fp32 FFMA path vs the forward + backward instance of the fused kernel
# (executed tensor flops per point: 6 forward + 6 backward 512 x 512 products)
for n_pts, prec in ((1 << 18, "fp32"), (1 << 20, "bf16"), (1 << 24, "bf16"), (1 << 24, "fp16")):
    pts = torch.rand((n_pts, 3), device=dev) * 2 - 1
    up = torch.randn(n_pts, device=dev) / n_pts
    ms = timed(lambda: dec.latent_vjp(z, pts, up, precision=prec))
    rows.append((f"latent_vjp {n_pts / 1e6:.2f}M points {prec} (forward + backward)", ms, n_pts / ms / 1e3,
                 0.0 if prec == "fp32" else n_pts * 2 * FLOP_Q / ms / 1e9))
    del pts, up
print(f"{'workload':58s} {'ms':>10s} {'M units/s':>12s} {'TFLOP/s':>9s}")
for name, ms, rate, tf in rows:
    print(f"{name:58s} {ms:10.3f} {rate:12.1f} {tf:9.1f}")
for n in (512, 4096):
    for mode in ("explicit", "seeded"):
        if mode == "explicit":
            g = torch.Generator(device=dev).manual_seed(1)
            x_T = torch.randn((n, 256), generator=g, device=dev)
            noise = torch.randn((1000, n, 256), generator=g, device=dev)
            fn = lambda: ddpm.sample_latents(n, x_T=x_T, noise=noise)
        else:
            fn = lambda: ddpm.sample_latents(n, seed=3)
        fn()
        ms = []
        for _ in range(3):
            fn()
            ms.append(ddpm.last_kernel_ms())
        m = statistics.median(ms)
        print(f"{'sample_latents(' + str(n) + ') 1000 steps bf16, ' + mode + ' noise':58s} {m:10.3f} {n / m / 1e3:12.4f} {n * 1000 * 7864320 / m / 1e9:9.1f}")
        if mode == "explicit":
            del noise
