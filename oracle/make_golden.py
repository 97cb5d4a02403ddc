"""Generate the committed fixtures under tests/golden/ from the CPU oracle.

    python -m oracle.make_golden            # rewrite tests/golden/oracle_golden.{npz,json}
    python -m oracle.make_golden --calibrate  # print the head bias that puts 25% of 64^3 inside
    python -m oracle.make_golden --vjp        # rewrite tests/golden/vjp_golden.npz only (latent-gradient fixtures)
    python -m oracle.make_golden --c-abi      # rewrite tests/golden/c_abi_decode32.bin (fp32 32^3 field for tests/c/gpu_decode.c)
    python -m oracle.make_golden --ddpm-rows  # rewrite tests/golden/ddpm_rows_golden.npz (64 rows of a seeded 4096-latent run)

There is no upstream reference to generate vectors from
(`/root/reference/README.md:1` is a title), so these fixtures pin the oracle to
itself: "frozen" becomes enforceable, and the GPU box (which has no
/root/reference either) checks the CUDA path against the same numbers.
Test infrastructure only.
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np
import torch

from . import (axis_coords, grid_points, sign_change_mask, decoder_weights, ddpm_weights,
               default_latent, weights_sha256, decode_grid, decoder_forward,
               decoder_forward_lowp, sample_latents, decoder_vjp_latent, decoder_vjp_latent_lowp, fit_loss_grad_lowp)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DDPM_NOISE_SEED = 2
DDPM_GOLDEN_N = 8


def ddpm_golden_inputs(n: int = DDPM_GOLDEN_N, steps: int = 1000):
    """x_T [n,256] and noise [steps,n,256] of the golden DDPM run (RandomState(2))."""
    rs = np.random.RandomState(DDPM_NOISE_SEED)
    x_T = rs.standard_normal((n, 256)).astype(np.float32)
    noise = rs.standard_normal((steps, n, 256)).astype(np.float32)
    return x_T, noise


def calibrate():
    params = [(w.copy(), b.copy()) for w, b in decoder_weights()]
    params[8] = (params[8][0], np.zeros(1, np.float32))
    sdf0 = decode_grid(default_latent(), 64, params=params)
    pre = np.arctanh(np.clip(sdf0.astype(np.float64), -0.999999, 0.999999))
    print("DEC_HEAD_BIAS =", repr(np.float32(-np.quantile(pre, 0.25))))


def vjp_golden_inputs():
    """Points, upstream gradient and fitting target of the latent-gradient fixtures (RandomState(13))."""
    rs = np.random.RandomState(13)
    xyz = (rs.rand(192, 3) * 2 - 1).astype(np.float32)
    up = ((0.5 + rs.rand(192)) / 192).astype(np.float32)          # coherent (all positive): no cancellation
    target = decoder_forward(default_latent(7), xyz)
    return xyz, up, target


def make_vjp_golden():
    """tests/golden/vjp_golden.npz: gradients of the SURVEY 8f row N4 oracles at the default latent (seed 2)."""
    torch.set_num_threads(8)
    z = default_latent(2)
    xyz, up, target = vjp_golden_inputs()
    arrays = {"xyz": xyz, "up": up, "target": target}
    arrays["grad_fp64"], arrays["sdf_fp64"] = decoder_vjp_latent(z, xyz, up)
    for name, lowp in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        arrays[f"grad_{name}"], arrays[f"sdf_{name}"] = decoder_vjp_latent_lowp(z, xyz, up, lowp=lowp)
        loss, g = fit_loss_grad_lowp(z, xyz, target, clamp=0.1, lowp=lowp)
        arrays[f"fit_loss_{name}"] = np.float64(loss)
        arrays[f"fit_grad_{name}"] = g
    np.savez_compressed(os.path.join(GOLDEN_DIR, "vjp_golden.npz"), **arrays)
    print({k: (v.shape, str(v.dtype)) for k, v in arrays.items()})


DDPM_ROWS_SEED = 77
DDPM_ROWS_N = 4096
DDPM_ROWS_BLOCKS = (0, 1000, 2555, 4080)     # four blocks of 16 consecutive latents of the 4096-latent batch


def ddpm_rows_inputs(steps: int = 1000):
    """(rows [64], x_T [64,256], noise [steps,64,256]): what the seeded sampler (seed DDPM_ROWS_SEED) draws for 64 of
    the 4096 latents of BASELINE configs[3] - the counters address latents by their global index, so the oracle can
    follow any subset of a batch (a latent's trajectory does not depend on the rest of the batch)."""
    from .philox import philox_normal_rows
    rows = np.concatenate([np.arange(b, b + 16) for b in DDPM_ROWS_BLOCKS])
    x_T = np.concatenate([philox_normal_rows(DDPM_ROWS_SEED, 16, steps, steps + 1, first_latent=b)[0] for b in DDPM_ROWS_BLOCKS])
    noise = np.concatenate([philox_normal_rows(DDPM_ROWS_SEED, 16, 0, steps, first_latent=b) for b in DDPM_ROWS_BLOCKS], axis=1)
    return rows, x_T, noise


def make_ddpm_rows_golden():
    torch.set_num_threads(os.cpu_count() or 1)
    rows, x_T, noise = ddpm_rows_inputs()
    arrays = {"rows": rows.astype(np.int64)}
    arrays["x0_fp32"] = sample_latents(len(rows), x_T, noise)
    arrays["x0_bf16"] = sample_latents(len(rows), x_T, noise, lowp=torch.bfloat16)
    arrays["x0_fp16"] = sample_latents(len(rows), x_T, noise, lowp=torch.float16)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ddpm_rows_golden.npz"), **arrays)
    print({k: (v.shape, str(v.dtype)) for k, v in arrays.items()},
          "bf16 vs fp32", float(np.abs(arrays["x0_bf16"] - arrays["x0_fp32"]).max()))


TRAIN_SAMPLE = 4096          # gradient entries kept per fixture (fixed indices, RandomState(5)), plus every tensor's norm


def train_golden_inputs():
    """Inputs of the training-step fixtures (second half): a DDPM batch of 48 latents and an auto-decoder
    batch of 2 shapes x 96 samples, all from RandomState(21)."""
    rs = np.random.RandomState(21)
    x0 = np.clip(rs.standard_normal((48, 256)) * 0.6, -1, 1).astype(np.float32)
    eps = rs.standard_normal((48, 256)).astype(np.float32)
    t = rs.randint(0, 1000, 48).astype(np.int32)
    lat = np.stack([default_latent(i) for i in range(2)])
    xyz = (rs.rand(2, 96, 3) * 2 - 1).astype(np.float32)
    tgt = np.stack([decoder_forward(default_latent(7 + i), xyz[i]) for i in range(2)])
    return (x0, t, eps), (lat, xyz, tgt)


def train_sample_indices(n_floats: int):
    return np.sort(np.random.RandomState(5).choice(n_floats, TRAIN_SAMPLE, replace=False))


def make_train_golden():
    """tests/golden/train_golden.npz: loss, sampled gradient entries and per-tensor gradient norms of the training-step
    oracles (oracle/train.py: fp64 autograd and the variants that emulate the 16-bit roundings of the tensor-core step)."""
    from .train import ddpm_train_grads, ddpm_train_grads_lowp, decoder_train_grads, decoder_train_grads_lowp, flatten_grads
    torch.set_num_threads(os.cpu_count() or 1)
    (x0, t, eps), (lat, xyz, tgt) = train_golden_inputs()
    arrays = {}

    def put(prefix, loss, grads):
        flat = flatten_grads(grads)
        arrays[prefix + "_loss"] = np.float64(loss)
        arrays[prefix + "_grad_sample"] = flat[train_sample_indices(flat.size)]
        arrays[prefix + "_grad_norm"] = np.float64(np.linalg.norm(flat.astype(np.float64)))
        arrays[prefix + "_tensor_norms"] = np.array([np.linalg.norm(np.asarray(g, dtype=np.float64)) for pair in grads for g in pair])

    l, g = ddpm_train_grads(x0, t, eps)
    put("ddpm_fp64", l, g)
    l, g, _ = decoder_train_grads(lat, xyz, tgt)
    put("dec_fp64", l, g)
    for name, lowp in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        l, g = ddpm_train_grads_lowp(x0, t, eps, lowp=lowp)
        put("ddpm_" + name, l, g)
        l, g, y = decoder_train_grads_lowp(lat, xyz, tgt, lowp=lowp)
        put("dec_" + name, l, g)
        arrays["dec_" + name + "_sdf"] = np.asarray(y, dtype=np.float32).ravel()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "train_golden.npz"), **arrays)
    print({k: (getattr(v, "shape", ()), str(getattr(v, "dtype", type(v)))) for k, v in arrays.items()})


def make_c_abi_golden():
    torch.set_num_threads(os.cpu_count() or 1)
    sdf = decode_grid(default_latent(), 32).astype(np.float32)
    sdf.tofile(os.path.join(GOLDEN_DIR, "c_abi_decode32.bin"))
    print("c_abi_decode32.bin", sdf.shape, float(sdf.min()), float(sdf.max()), int((sdf < 0).sum()), "inside")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calibrate", action="store_true")
    ap.add_argument("--vjp", action="store_true")
    ap.add_argument("--c-abi", action="store_true")
    ap.add_argument("--ddpm-rows", action="store_true")
    ap.add_argument("--train", action="store_true")
    args = ap.parse_args()
    if args.train:
        return make_train_golden()
    if args.calibrate:
        return calibrate()
    if args.vjp:
        return make_vjp_golden()
    if args.c_abi:
        return make_c_abi_golden()
    if args.ddpm_rows:
        return make_ddpm_rows_golden()
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    z = default_latent()
    arrays, meta = {}, {}
    meta["decoder_sha256"] = weights_sha256(decoder_weights())
    meta["ddpm_sha256"] = weights_sha256(ddpm_weights())
    meta["latent_sha256"] = __import__("hashlib").sha256(z.tobytes()).hexdigest()
    for res in (64, 128, 256, 512):
        arrays[f"coords_{res}"] = axis_coords(res)

    sdf64 = decode_grid(z, 64)
    mask64 = sign_change_mask(sdf64)
    meta["sdf64_inside"] = int((sdf64 < 0).sum())
    meta["sdf64_active_cells"] = int(mask64.sum())
    meta["sdf64_min"] = float(sdf64.min())
    meta["sdf64_max"] = float(sdf64.max())
    arrays["mask64_bits"] = np.packbits(mask64.ravel())
    rs = np.random.RandomState(7)
    idx = np.sort(rs.choice(64 ** 3, 512, replace=False)).astype(np.int64)
    pts64 = grid_points(64)
    arrays["sdf64_idx"] = idx
    arrays["sdf64_fp32"] = sdf64.ravel()[idx]
    arrays["sdf64_bf16"] = decoder_forward_lowp(z, pts64[idx], lowp=torch.bfloat16)
    arrays["sdf64_fp16"] = decoder_forward_lowp(z, pts64[idx], lowp=torch.float16)
    # a z-slab the multi-GPU tests decode on their own: planes [16, 24) of 64^3
    arrays["sdf64_slab_16_24"] = sdf64[16:24].copy()
    # sparse samples of the big grids (the full 256^3 / 512^3 oracle is minutes of CPU)
    for res in (256, 512):
        c = axis_coords(res)
        q = np.sort(rs.choice(res ** 3, 256, replace=False)).astype(np.int64)
        ix, iy, iz = q % res, (q // res) % res, q // (res * res)
        pts = np.stack([c[ix], c[iy], c[iz]], axis=1)
        arrays[f"sdf{res}_idx"] = q
        arrays[f"sdf{res}_fp32"] = decoder_forward(z, pts)
        arrays[f"sdf{res}_bf16"] = decoder_forward_lowp(z, pts, lowp=torch.bfloat16)
    # a second latent, arbitrary (off-grid) points: the Decoder(latent, xyz) entry
    z1 = default_latent(1)
    pts = (rs.uniform(-1, 1, size=(256, 3))).astype(np.float32)
    arrays["points_xyz"] = pts
    arrays["points_fp32"] = decoder_forward(z1, pts)
    arrays["points_bf16"] = decoder_forward_lowp(z1, pts, lowp=torch.bfloat16)

    x_T, noise = ddpm_golden_inputs()
    arrays["ddpm_fp32"] = sample_latents(DDPM_GOLDEN_N, x_T, noise)
    arrays["ddpm_bf16"] = sample_latents(DDPM_GOLDEN_N, x_T, noise, lowp=torch.bfloat16)
    x64 = sample_latents(DDPM_GOLDEN_N, x_T, noise, dtype=torch.float64)
    meta["ddpm_fp32_vs_fp64_maxabs"] = float(np.abs(arrays["ddpm_fp32"] - x64).max())
    sdf64_f64 = decode_grid(z, 64, dtype=torch.float64)
    meta["sdf64_fp32_vs_fp64_maxabs"] = float(np.abs(sdf64 - sdf64_f64).max())
    meta["torch"] = torch.__version__
    meta["numpy"] = np.__version__

    np.savez_compressed(os.path.join(GOLDEN_DIR, "oracle_golden.npz"), **arrays)
    with open(os.path.join(GOLDEN_DIR, "oracle_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
