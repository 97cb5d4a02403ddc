"""Multi-rank GPU check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py [res]

Config 5 (one latent, res^3 grid, z-slab per rank, fused mask with locally recomputed halo plane,
in-place NCCL all-gather) and config 3 (a batch of latents split across ranks) are compared bit for bit
with what ONE GPU computes on its own.  Prints one JSON line on rank 0; exit code 0 = pass.
Launched by tests/test_gpu_multirank.py when >= 2 GPUs are visible."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from __graft_entry__ import load_package
import oracle


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = load_package()
    dec = pkg.Decoder(oracle.flatten_params(oracle.decoder_weights()), device=dev, precision="bf16")
    z = torch.from_numpy(oracle.default_latent()).to(dev)

    # config 5: z-slabs + mask + all-gather, timed on the device (max over ranks)
    for _ in range(2):
        sdf, mask = pkg.decode_grid_sharded(dec, z, res, mask=True)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sdf, mask = pkg.decode_grid_sharded(dec, z, res, mask=True)
    e1.record(); e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref_sdf, ref_mask = dec.decode_grid(z, res, mask=True)           # the whole grid on this GPU alone
    ok_sdf = bool(torch.equal(sdf, ref_sdf))
    ok_mask = bool(torch.equal(mask, ref_mask))

    # config 3: a batch of latents split across ranks, no communication
    B, bres = 2 * world + 1, 32
    lat = torch.stack([torch.from_numpy(oracle.default_latent(i)) for i in range(B)]).to(dev)
    i0, part = pkg.decode_batch_sharded(dec, lat, bres)
    ok_batch = all(bool(torch.equal(part[k], dec.decode_grid(lat[i0 + k], bres))) for k in range(part.shape[0]))
    # config 4's sampling phase: shares of one seeded batch sampled per rank == the whole batch sampled by one GPU
    ddpm = pkg.LatentDDPM(oracle.flatten_params(oracle.ddpm_weights()), device=dev, precision="bf16")
    os.environ["SDFB_DDPM_BN"] = "128"          # the tile width fixes the fp32 summation order: same for shares and whole
    n_lat, steps = 150 * world + 7, 12
    gathered = pkg.sample_latents_sharded(ddpm, n_lat, seed=31, steps=steps, gather=True)
    whole = ddpm.sample_latents(n_lat, steps=steps, seed=31)
    ok_sample = bool(torch.equal(gathered, whole))
    ok_batch = ok_batch and ok_sample

    counts = torch.tensor([part.shape[0]], device=dev)
    dist.all_reduce(counts)
    ok = torch.tensor([int(ok_sdf and ok_mask and ok_batch and int(counts.item()) == B)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "res": res, "sdf_equal": ok_sdf, "mask_equal": ok_mask, "batch_equal": ok_batch, "sharded_sampling_equal": ok_sample,
                          "all_ranks_ok": bool(ok.item()), "sharded_decode_mask_gather_ms": float(ms.item()),
                          "queries_per_s": res ** 3 / (float(ms.item()) * 1e-3)}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if bool(ok.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
