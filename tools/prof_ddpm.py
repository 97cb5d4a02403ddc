#!/usr/bin/env python
"""Run the fused DDPM sampler once per tile width and print device time (and, with SDFB_PROF=1,
the blocked-cycle profile per role).  usage: python tools/prof_ddpm.py [n] [steps] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pkg = load_package()
m = pkg.LatentDDPM(pkg.synthetic.ddpm_params(), device="cuda:0", precision="bf16")
g = torch.Generator(device="cuda").manual_seed(0)
x_T = torch.randn((n, 256), generator=g, device="cuda")
noise = torch.randn((steps, n, 256), generator=g, device="cuda")
for r in range(reps):
    m.sample_latents(n, x_T=x_T, noise=noise, steps=steps)
    ms = m.last_kernel_ms()
    print(f"n={n} steps={steps} bn={os.environ.get('SDFB_DDPM_BN', 'auto')}: {ms:.3f} ms = {ms * 1e3 / steps:.2f} us/step, "
          f"{n * steps * 7864320 / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)
