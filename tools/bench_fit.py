#!/usr/bin/env python
"""Dev tool: wall time of Decoder.fit_latent (auto-decoder inference) per precision: n points, `steps` Adam steps."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
pkg = load_package()
dec = pkg.Decoder(pkg.synthetic.decoder_params())
z_true = torch.from_numpy(pkg.synthetic.latent(7)).cuda()
pts = torch.rand((n, 3), device="cuda") * 2 - 1
tgt = dec(z_true, pts, precision="fp32")
held = torch.rand((20000, 3), device="cuda") * 2 - 1
want = torch.clamp(dec(z_true, held, precision="fp32"), -0.1, 0.1)
for prec, st in (("bf16", steps), ("fp16", steps), ("fp32", max(steps // 10, 2))):
    dec.fit_latent(pts, tgt, steps=2, precision=prec)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    z, loss = dec.fit_latent(pts, tgt, steps=st, lr=1e-2, reg=0.0, precision=prec)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = float((torch.clamp(dec(z, held, precision="fp32"), -0.1, 0.1) - want).abs().mean())
    print(f"fit_latent {prec}: {n} points x {st} steps in {dt * 1e3:.1f} ms = {dt / st * 1e3:.3f} ms/step "
          f"({n * st / dt / 1e6:.1f} M point-steps/s); train loss {loss:.5f}, held-out clamped-L1 {err:.5f}")
