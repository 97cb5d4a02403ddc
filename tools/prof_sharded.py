#!/usr/bin/env python
"""Dev tool (run under torchrun, one rank per GPU): where does the time of sdfb_decode_grid_sharded go?
Times the 512^3 sharded decode for a few sub-slab sizes, with and without the mask, with the pushes switched off
(SDFB_PUSH_OFF=1: results on the peers are then invalid) and against a plain decode of the rank's slab."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pkg = load_package()
res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dec = pkg.Decoder(pkg.synthetic.decoder_params(), device=dev)
z = torch.from_numpy(pkg.synthetic.latent(0)).to(dev)
comm = pkg.Comm(dev)


def timed(fn, reps=4):
    fn(); fn()
    ms = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    t = torch.tensor([statistics.median(ms)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


z0, z1 = pkg.slab_range(res, rank, world)
rows = []
rows.append(("plain decode of own slab + mask (one launch)", timed(lambda: dec.decode_grid(z, res, z0, z1, mask=True))))
rows.append(("plain decode of own slab, no mask", timed(lambda: dec.decode_grid(z, res, z0, z1))))
for sub in (0, 16, 32, 64, 128, 512):
    rows.append((f"sharded push path, mask, sub_planes={sub}", timed(lambda: comm.decode_grid_sharded(dec, z, res, mask=True, sub_planes=sub))))
rows.append(("sharded push path, no mask, sub_planes=0", timed(lambda: comm.decode_grid_sharded(dec, z, res, mask=False))))
os.environ["SDFB_PUSH_OFF"] = "1"
for sub in (0, 64):
    rows.append((f"sharded, pushes OFF, mask, sub_planes={sub}", timed(lambda: comm.decode_grid_sharded(dec, z, res, mask=True, sub_planes=sub))))
os.environ["SDFB_PUSH_OFF"] = "2"
rows.append(("sharded, pushes + barriers OFF, mask, sub_planes=0", timed(lambda: comm.decode_grid_sharded(dec, z, res, mask=True))))
del os.environ["SDFB_PUSH_OFF"]
rows.append(("torch.distributed path (decode, NCCL all-gathers)", timed(lambda: pkg.decode_grid_sharded(dec, z, res, mask=True))))
if rank == 0:
    for name, ms in rows:
        print(f"{name:60s} {ms:9.3f} ms")
dist.barrier()
dist.destroy_process_group()
