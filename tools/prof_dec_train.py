import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from __graft_entry__ import load_package
pkg = load_package()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, P = 64, 8192
dt = pkg.DecoderTrainer(pkg.synthetic.decoder_params(), device=dev)
lat = torch.stack([torch.from_numpy(pkg.synthetic.latent(i % 8)) for i in range(B)]).to(dev)
xyz = torch.rand((B, P, 3), generator=g, device=dev) * 2 - 1
tgt = torch.rand((B, P), generator=g, device=dev) * 0.2 - 0.1
for _ in range(2):
    loss = dt.step(lat, xyz, tgt, lr=1e-5)
print("decoder train loss", float(loss.item()))
