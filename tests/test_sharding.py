"""CPU tests of the multi-GPU partition logic, including a world_size-2 gloo run in which a
stand-in decoder (the oracle) takes the place of the CUDA one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def test_slab_range_covers_grid(pkg):
    for res in (2, 63, 64, 128, 512):
        for world in (1, 2, 3, 4, 8):
            planes = []
            for r in range(world):
                z0, z1 = pkg.slab_range(res, r, world)
                assert 0 <= z0 <= z1 <= res
                planes += list(range(z0, z1))
            assert planes == list(range(res))
    assert pkg.slab_range(512, 3, 8) == (192, 256)
    with pytest.raises(ValueError):
        pkg.slab_range(64, 2, 2)


def test_batch_range_balanced(pkg):
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 4, 8):
            items, sizes = [], []
            for r in range(world):
                i0, i1 = pkg.batch_range(n, r, world)
                items += list(range(i0, i1))
                sizes.append(i1 - i0)
            assert items == list(range(n))
            assert max(sizes) - min(sizes) <= 1


class _OracleDecoder:
    """Stands in for pkg.Decoder on CPU: same decode_grid contract (slab + halo + mask)."""
    device = torch.device("cpu")

    def __init__(self, field):
        self.field = field        # precomputed full sdf [res,res,res]

    def decode_grid(self, latent, res, z0=0, z1=None, mask=False, precision=None, out=None):
        z1 = res if z1 is None else z1
        halo = 1 if (mask and z1 < res and z1 > z0) else 0
        sdf = torch.from_numpy(self.field[z0:z1 + halo].copy())
        if not mask:
            return sdf[: z1 - z0]
        m = torch.from_numpy(oracle.sign_change_mask(sdf.numpy())) if sdf.shape[0] >= 2 else torch.zeros((0, res - 1, res - 1), dtype=torch.uint8)
        return sdf[: z1 - z0], m


class _OracleSampler:
    """Stands in for pkg.LatentDDPM on CPU: seeded sampling addressed by global latent index."""
    device = torch.device("cpu")

    def sample_latents(self, n, steps=1000, seed=0, precision=None, first_latent=0):
        noise = oracle.philox_normal_rows(seed, n, 0, steps, first_latent=first_latent)
        x_T = oracle.philox_normal_rows(seed, n, steps, steps + 1, first_latent=first_latent)[0]
        return torch.from_numpy(oracle.sample_latents(n, x_T, noise, steps=steps))


class _StandInFitter:
    """Stands in for pkg.Decoder.fit_latents_batch on CPU: a deterministic function of each shape's own samples."""
    device = torch.device("cpu")

    @staticmethod
    def fit_latents_batch(xyz, sdf_target, **kw):
        z = torch.stack([torch.full((256,), float(t.mean())) + x.sum() for x, t in zip(xyz, sdf_target)])
        return z, sdf_target.abs().mean(dim=1)


def _worker(rank, world, port, res, field, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        pkg = importlib.import_module("latent-diffusion-models-for-shape-sdfs_b200")
        dec = _OracleDecoder(field)
        sdf, m = pkg.decode_grid_sharded(dec, None, res, mask=True)
        ok = bool(np.array_equal(sdf.numpy(), field)) and bool(np.array_equal(m.numpy(), oracle.sign_change_mask(field)))
        i0, i1 = pkg.batch_range(5, rank, world)
        # sharded sampling: a stand-in sampler (oracle sampler on the oracle's Philox stream, 3 steps) per rank;
        # the gathered batch must equal one process sampling all latents
        full = pkg.sample_latents_sharded(_OracleSampler(), 5, seed=77, steps=3, gather=True)
        x_T, noise = oracle.philox_sampler_inputs(77, 5, 3)
        ok = ok and bool(np.array_equal(full.numpy(), oracle.sample_latents(5, x_T, noise, steps=3)))
        # sharded fitting of 5 shapes: the gathered latents / losses equal one process fitting all of them
        g = torch.Generator().manual_seed(3)
        xyz, tgt = torch.rand((5, 7, 3), generator=g), torch.rand((5, 7), generator=g)
        zs, ls = pkg.fit_latents_sharded(_StandInFitter(), xyz, tgt, gather=True)
        z_all, l_all = _StandInFitter.fit_latents_batch(xyz, tgt)
        ok = ok and bool(torch.equal(zs, z_all)) and bool(torch.equal(ls, l_all))
        j0, zl, _ = pkg.fit_latents_sharded(_StandInFitter(), xyz, tgt)
        ok = ok and j0 == i0 and bool(torch.equal(zl, z_all[i0:i1]))
        out_q.put((rank, ok, (i0, i1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("res", [12, 13])
def test_sharded_decode_gloo_world2(res):
    # an analytic field with a surface crossing slab boundaries (sphere of radius 0.6)
    c = oracle.axis_coords(res)
    zz, yy, xx = np.meshgrid(c, c, c, indexing="ij")
    field = (np.sqrt(xx * xx + yy * yy + zz * zz) - 0.6).astype(np.float32)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, res, field, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    assert sorted(r[2] for r in results) == [(0, 3), (3, 5)]


def test_partition_properties_hypothesis(pkg):
    """slab_range / batch_range: disjoint, ordered, covering, balanced - for arbitrary sizes."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(2, 4096), st.integers(1, 64))
    def slabs(res, world):
        prev = 0
        for r in range(world):
            z0, z1 = pkg.slab_range(res, r, world)
            assert z0 == min(prev, res) and z0 <= z1 <= res
            prev = z1
        assert prev == res

    @settings(max_examples=200, deadline=None)
    @given(st.integers(0, 100000), st.integers(1, 64))
    def batches(n, world):
        prev, sizes = 0, []
        for r in range(world):
            i0, i1 = pkg.batch_range(n, r, world)
            assert i0 == prev and i1 >= i0
            prev = i1
            sizes.append(i1 - i0)
        assert prev == n and max(sizes) - min(sizes) <= 1

    slabs()
    batches()
