#!/usr/bin/env python
"""Dev tool: a few training steps (DDPM denoiser and auto-decoder) for profiling the general tensor-core product."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package

pkg = load_package()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tr = pkg.DDPMTrainer(pkg.synthetic.ddpm_params(), device=dev)
x0 = torch.randn((n, 256), generator=g, device=dev).clamp_(-1, 1)
eps = torch.randn((n, 256), generator=g, device=dev)
t = torch.randint(0, 1000, (n,), generator=g, device=dev, dtype=torch.int32)
for _ in range(3):
    loss = tr.step(x0, t, eps)
print("ddpm train loss", float(loss.item()))
B, P = 16, 8192
dt = pkg.DecoderTrainer(pkg.synthetic.decoder_params(), device=dev)
lat = torch.stack([torch.from_numpy(pkg.synthetic.latent(i)) for i in range(B)]).to(dev)
xyz = torch.rand((B, P, 3), generator=g, device=dev) * 2 - 1
tgt = torch.rand((B, P), generator=g, device=dev) * 0.2 - 0.1
for _ in range(2):
    loss = dt.step(lat, xyz, tgt, lr=1e-5)
print("decoder train loss", float(loss.item()))
