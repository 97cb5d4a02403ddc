"""Multi-GPU partitioning: z-slabs of one grid, or independent
latents.  One process per GPU; the only collective is the final all-gather of the slabs
(NCCL over NVLink on the box, gloo in the CPU tests).  Index math here is pure Python so
it is testable without a GPU."""
from __future__ import annotations

import torch
import torch.distributed as dist


def slab_range(res: int, rank: int, world: int) -> tuple[int, int]:
    """Planes [z0, z1) owned by ``rank``: ceil(res/world) planes each, the tail ranks may be
    short or empty.  Equal-sized when world divides res (the all-gather is then in place)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-res // world)
    z0 = min(rank * per, res)
    return z0, min(z0 + per, res)


def batch_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Items [i0, i1) of a batch of n independent latents owned by ``rank`` (balanced +-1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    i0 = rank * base + min(rank, extra)
    return i0, i0 + base + (1 if rank < extra else 0)


def gather_slabs(local: torch.Tensor, res_planes: int, per: int, group=None) -> torch.Tensor:
    """All-gather equal-sized (padded to ``per`` leading rows) slabs and cut to ``res_planes``."""
    world = dist.get_world_size(group)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    return full[:res_planes]


def decode_grid_sharded(decoder, latent, res: int, mask=False, precision=None, group=None,
                        gather: bool = True, comm=None):
    """Config 5: every rank decodes its z-slab (plus, for the mask, the halo plane above it,
    recomputed locally - no halo exchange), then the slabs are assembled on every rank.

    Returns (sdf, mask_or_None); full [res,res,res] / [(res-1)^3] tensors if ``gather`` else the
    rank's own slab.

    ``comm`` (a ``Comm``): the overlapped path - one C call (``sdfb_decode_grid_sharded``) that pushes every finished
    sub-slab into all peers' copies of a symmetric buffer with the copy engines while the next one is being decoded; the
    results are views of that buffer, valid until the next sharded decode.  ``mask="bits"`` then returns the packed mask
    blocks [world, words_per_rank] as they travelled (``unpack_mask_blocks`` gives the uint8 form ``mask=True`` returns).
    Without ``comm``: decode, then ``torch.distributed`` all-gathers (NCCL on the box, gloo in the CPU tests)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if comm is not None and gather:
        sdf_full, words = comm.decode_grid_sharded(decoder, latent, res, mask=bool(mask), precision=precision)
        if not mask:
            return sdf_full, None
        if mask == "bits":
            return sdf_full, words
        from .api import unpack_mask_blocks
        return sdf_full, unpack_mask_blocks(words, res)
    if mask == "bits":
        raise ValueError('mask="bits" needs the overlapped path (pass comm=)')
    z0, z1 = slab_range(res, rank, world)
    per = -(-res // world)
    even = res % world == 0
    if even and gather and decoder.device.type == "cuda":
        # decode straight into this rank's block of the full buffer; NCCL gathers in place
        # (a rank's halo plane lands in its neighbour's block and is then overwritten by the
        # neighbour's identical values)
        full = torch.empty((res, res, res), dtype=torch.float32, device=decoder.device)
        flat = full.view(-1)[z0 * res * res:]
        r = decoder.decode_grid(latent, res, z0, z1, mask=mask, precision=precision, out=flat)
        sdf_local, m_local = (r if mask else (r, None))
        sdf_full = full
        dist.all_gather_into_tensor(sdf_full, sdf_full[z0:z1], group=group)
    else:
        r = decoder.decode_grid(latent, res, z0, z1, mask=mask, precision=precision)
        sdf_local, m_local = (r if mask else (r, None))
        if not gather:
            return sdf_local, m_local
        sdf_full = gather_slabs(sdf_local, res, per, group)
    m_full = None
    if mask:
        m_full = gather_slabs(m_local, res - 1, per, group)
    return sdf_full, m_full


def decode_batch_sharded(decoder, latents: torch.Tensor, res: int, precision=None, group=None):
    """Configs 3/4: latents [B,256] are split across ranks; each rank decodes its share.
    No communication; returns (i0, sdf [i1-i0, res, res, res])."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    i0, i1 = batch_range(latents.shape[0], rank, world)
    if hasattr(decoder, "decode_grid_batch"):
        return i0, decoder.decode_grid_batch(latents[i0:i1], res, precision=precision)
    out = torch.empty((i1 - i0, res, res, res), dtype=torch.float32, device=decoder.device)
    for j, i in enumerate(range(i0, i1)):
        decoder.decode_grid(latents[i], res, precision=precision, out=out[j])
    return i0, out


def sample_latents_sharded(sampler, n: int, seed: int = 0, steps: int = 1000, precision=None, group=None, gather: bool = False):
    """Config 4's sampling phase: the batch of n latents is split across ranks; every rank samples its share with
    the in-kernel noise addressed by GLOBAL latent index, so every rank draws exactly the noise a single call over the
    whole batch would.  The samples themselves equal the whole-batch ones up to fp32 summation order: the fused kernel
    picks its tile width from the per-call latent count, and a different width sums a layer's products in a different
    order (bit-equal when the widths coincide, e.g. with SDFB_DDPM_BN set; the fp32 path is always bit-equal).
    No communication unless ``gather``.  Returns (i0, x [i1-i0, 256]) or the gathered [n, 256]."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    i0, i1 = batch_range(n, rank, world)
    if i1 > i0:
        x = sampler.sample_latents(i1 - i0, steps=steps, seed=seed, precision=precision, first_latent=i0)
    else:
        x = torch.empty((0, 256), dtype=torch.float32, device=sampler.device)
    if not gather:
        return i0, x
    per = -(-n // world)
    pad = torch.zeros((per, 256), dtype=torch.float32, device=x.device)
    pad[: x.shape[0]] = x
    full = torch.empty((world * per, 256), dtype=torch.float32, device=x.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    sizes = [batch_range(n, r, world) for r in range(world)]
    return torch.cat([full[r * per: r * per + (b - a)] for r, (a, b) in enumerate(sizes)])


def fit_latents_sharded(decoder, xyz: torch.Tensor, sdf_target: torch.Tensor, group=None, gather: bool = False, **fit_kwargs):
    """Auto-decoder fitting of a batch of shapes, shapes split across ranks: xyz [B,M,3],
    sdf_target [B,M]; every rank fits its share with ``decoder.fit_latents_batch``.  Shapes are independent, so there is
    no communication unless ``gather``.  Returns (i0, latents [i1-i0,256], losses [i1-i0]) or the gathered
    (latents [B,256], losses [B])."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = xyz.shape[0]
    i0, i1 = batch_range(B, rank, world)
    if i1 > i0:
        z, loss = decoder.fit_latents_batch(xyz[i0:i1], sdf_target[i0:i1], **fit_kwargs)
    else:
        z = torch.empty((0, 256), dtype=torch.float32, device=decoder.device)
        loss = torch.empty((0,), dtype=torch.float32, device=decoder.device)
    if not gather:
        return i0, z, loss
    per = -(-B // world)
    pad = torch.zeros((per, 257), dtype=torch.float32, device=z.device)      # latent | loss
    pad[: z.shape[0], :256] = z
    pad[: z.shape[0], 256] = loss
    full = torch.empty((world * per, 257), dtype=torch.float32, device=z.device)
    dist.all_gather_into_tensor(full, pad, group=group)
    sizes = [batch_range(B, r, world) for r in range(world)]
    out = torch.cat([full[r * per: r * per + (b - a)] for r, (a, b) in enumerate(sizes)])
    return out[:, :256].contiguous(), out[:, 256].contiguous()
