#!/usr/bin/env python
"""Dev tool: time the forward + backward (latent gradient) instance of the fused decoder kernel on n random points
(and, with SDFB_PROF=1 in the environment, print the per-role blocked-cycle profile)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pkg = load_package()
dec = pkg.Decoder(pkg.synthetic.decoder_params(), precision=prec)
z = torch.from_numpy(pkg.synthetic.latent(0)).cuda()
pts = torch.rand((n, 3), device="cuda") * 2 - 1
up = torch.randn(n, device="cuda") / n
for i in range(reps):
    dec.latent_vjp(z, pts, up, precision=prec)
    torch.cuda.synchronize()
    ms = dec.last_kernel_ms()
    print(f"latent_vjp {n} points {prec}: {ms:.3f} ms  {n / ms / 1e3:.1f} M points/s  "
          f"{n * 2 * 3145728 / ms / 1e9:.1f} TFLOP/s (tensor-pipe flops, forward + backward)")
