#!/usr/bin/env python
"""Dev tool: decode one res^3 grid a few times and print the fused kernel's time (and, with
SDFB_PROF=1 in the environment, the per-role blocked-cycle profile)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pkg = load_package()
dec = pkg.Decoder(pkg.synthetic.decoder_params(), precision=prec)
z = torch.from_numpy(pkg.synthetic.latent(0)).cuda()
out = torch.empty((res, res, res), device="cuda")
for i in range(reps):
    dec.decode_grid(z, res, out=out)
    torch.cuda.synchronize()
    ms = dec.last_kernel_ms()
    q = res ** 3
    print(f"{res}^3 {prec}: {ms:.3f} ms  {q / ms / 1e6:.1f} Gq/s  {q * 3145728 / ms / 1e9:.1f} TFLOP/s(tensor-pipe flops)")
