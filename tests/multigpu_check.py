"""Multi-rank GPU check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py [res]

Config 5 (one latent, res^3 grid, z-slab per rank, fused mask with locally recomputed halo plane,
in-place NCCL all-gather; and the overlapped peer-memory path of sdfb_decode_grid_sharded, even and uneven splits) and
config 3 (a batch of latents split across ranks) are compared bit for bit with what ONE GPU computes on its own.  Prints one JSON line on rank 0; exit code 0 = pass.
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

    # the overlapped path: ONE C call per rank (sdfb_decode_grid_sharded) - sub-slabs pushed into every peer's copy of a
    # symmetric buffer by the copy engines while the next one is decoded, the mask travelling packed
    comm = pkg.Comm(dev)
    for _ in range(2):
        sdf2, words = comm.decode_grid_sharded(dec, z, res, mask=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    sdf2, words = comm.decode_grid_sharded(dec, z, res, mask=True)
    e1.record(); e1.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ok_push = bool(torch.equal(sdf2, ref_sdf)) and bool(torch.equal(pkg.unpack_mask_blocks(words, res), ref_mask))
    # uneven splits, including ranks whose slab is empty (res 33 over 8 ranks: per = 5, rank 7 gets [33, 33))
    for r2 in (33, 50):
        s3, w3 = comm.decode_grid_sharded(dec, z, r2, mask=True)
        rs3, rm3 = dec.decode_grid(z, r2, mask=True)
        ok_push = ok_push and bool(torch.equal(s3, rs3)) and bool(torch.equal(pkg.unpack_mask_blocks(w3, r2), rm3))
        s4, _ = pkg.decode_grid_sharded(dec, z, r2, mask=True)       # torch.distributed path on the same uneven split
        ok_push = ok_push and bool(torch.equal(s4, rs3))
    # the C-ABI in-place all-gather (sdfb_allgather_slabs) on an even split
    full = torch.zeros((8 * world, 1024), dtype=torch.float32, device=dev)
    full[8 * rank: 8 * rank + 8] = float(rank + 1)
    comm.allgather_slabs(full)
    comm.barrier()
    torch.cuda.synchronize()
    want = torch.arange(1, world + 1, device=dev, dtype=torch.float32).repeat_interleave(8)[:, None].expand(-1, 1024)
    ok_push = ok_push and bool(torch.equal(full, want))
    dec.check()

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
    ok = torch.tensor([int(ok_sdf and ok_mask and ok_batch and ok_push and int(counts.item()) == B)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "res": res, "sdf_equal": ok_sdf, "mask_equal": ok_mask, "batch_equal": ok_batch, "sharded_sampling_equal": ok_sample,
                          "push_path_equal": ok_push, "push_path_ms": float(ms2.item()),
                          "all_ranks_ok": bool(ok.item()), "sharded_decode_mask_gather_ms": float(ms.item()),
                          "queries_per_s": res ** 3 / (float(ms.item()) * 1e-3)}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if bool(ok.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
