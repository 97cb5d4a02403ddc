"""CPU tests: the oracle against its committed golden fixtures and its own fp64
re-evaluation (there are no upstream vectors: /root/reference/README.md:1 is a title)."""
import hashlib

import numpy as np
import torch

import oracle
from oracle.make_golden import ddpm_golden_inputs


def test_weights_are_frozen(golden):
    _, meta = golden
    assert oracle.weights_sha256(oracle.decoder_weights()) == meta["decoder_sha256"]
    assert oracle.weights_sha256(oracle.ddpm_weights()) == meta["ddpm_sha256"]
    assert hashlib.sha256(oracle.default_latent().tobytes()).hexdigest() == meta["latent_sha256"]
    flat = oracle.flatten_params(oracle.decoder_weights())
    assert flat.size == 1_835_520 + 3_838


def test_axis_coords_bit_exact(golden):
    arrays, _ = golden
    for res in (64, 128, 256, 512):
        c = oracle.axis_coords(res)
        assert c.dtype == np.float32
        assert np.array_equal(c.view(np.uint32), arrays[f"coords_{res}"].view(np.uint32))
        assert c[0] == -1.0 and c[-1] == 1.0
        # antisymmetric by construction
        assert np.array_equal(c, -c[::-1])


def test_grid_points_order():
    p = oracle.grid_points(4)
    c = oracle.axis_coords(4)
    assert p.shape == (64, 3)
    q = (2 * 4 + 1) * 4 + 3          # iz=2, iy=1, ix=3
    assert tuple(p[q]) == (c[3], c[1], c[2])
    slab = oracle.grid_points(4, 1, 3)
    assert np.array_equal(slab, p[16:48])


def test_mask_definition():
    s = np.ones((3, 3, 3), np.float32)
    assert oracle.sign_change_mask(s).sum() == 0
    s[1, 1, 1] = -1.0                       # centre node touches all 8 cells
    assert oracle.sign_change_mask(s).sum() == 8
    s[:] = -1.0
    assert oracle.sign_change_mask(s).sum() == 0
    s = np.ones((2, 2, 2), np.float32)
    s[0, 0, 0] = -0.0                       # -0 is outside
    assert oracle.sign_change_mask(s).sum() == 0
    s[0, 0, 0] = np.nan                     # NaN is outside
    assert oracle.sign_change_mask(s).sum() == 0


def test_decoder_golden_64(golden):
    arrays, meta = golden
    torch.set_num_threads(8)
    sdf = oracle.decode_grid(oracle.default_latent(), 64)
    assert sdf.shape == (64, 64, 64) and sdf.dtype == np.float32
    np.testing.assert_allclose(sdf.ravel()[arrays["sdf64_idx"]], arrays["sdf64_fp32"], atol=2e-6, rtol=0)
    # non-degenerate field: a real surface inside the cube
    assert abs(int((sdf < 0).sum()) - meta["sdf64_inside"]) <= 8
    mask = oracle.sign_change_mask(sdf)
    golden_mask = np.unpackbits(arrays["mask64_bits"])[:63 ** 3].reshape(63, 63, 63)
    assert (mask != golden_mask).sum() <= 16     # sgemm summation order may flip |sdf|<1e-6 nodes
    assert meta["sdf64_active_cells"] > 5000
    np.testing.assert_allclose(sdf[16:24], arrays["sdf64_slab_16_24"], atol=2e-6, rtol=0)


def test_decoder_fp32_vs_fp64_noise_floor(golden):
    arrays, _ = golden
    z = oracle.default_latent()
    pts = oracle.grid_points(64)[arrays["sdf64_idx"]]
    f32 = oracle.decoder_forward(z, pts)
    f64 = oracle.decoder_forward(z, pts, dtype=torch.float64)
    assert np.abs(f32 - f64).max() < 5e-6       # the 1e-5 fp32 criterion sits above this floor


def test_decoder_lowp_golden_and_distance(golden):
    arrays, _ = golden
    z = oracle.default_latent()
    pts = oracle.grid_points(64)[arrays["sdf64_idx"]]
    bf = oracle.decoder_forward_lowp(z, pts, lowp=torch.bfloat16)
    np.testing.assert_allclose(bf, arrays["sdf64_bf16"], atol=5e-4, rtol=0)
    fh = oracle.decoder_forward_lowp(z, pts, lowp=torch.float16)
    np.testing.assert_allclose(fh, arrays["sdf64_fp16"], atol=1e-4, rtol=0)
    ref = arrays["sdf64_fp32"]
    # measured distances to the fp32 oracle: bf16 misses 2e-3, fp16 meets it
    assert 2e-3 < np.abs(bf - ref).max() < 2e-2
    assert np.abs(fh - ref).max() < 2e-3
    m = np.abs(ref) > 2e-3
    assert ((bf < 0) == (ref < 0))[m].mean() >= 0.999


def test_decoder_batched_latents_and_points(golden):
    arrays, _ = golden
    z1 = oracle.default_latent(1)
    out = oracle.decoder_forward(z1, arrays["points_xyz"])
    np.testing.assert_allclose(out, arrays["points_fp32"], atol=2e-6, rtol=0)
    zz = np.stack([z1] * 4)
    out_b = oracle.decoder_forward(zz, arrays["points_xyz"][:4])
    np.testing.assert_allclose(out_b, arrays["points_fp32"][:4], atol=2e-6, rtol=0)


def test_ddpm_schedule_and_embedding():
    s = oracle.ddpm_schedule()
    assert all(v.shape == (1000,) and v.dtype == np.float32 for v in s.values())
    assert s["sigma"][0] == 0.0 and s["c2"][0] == 0.0
    np.testing.assert_allclose(s["c1"][0], 1.0, rtol=1e-6)
    te = oracle.time_embedding([0, 999])
    assert te.shape == (2, 256)
    assert np.all(te[0, :128] == 0) and np.all(te[0, 128:] == 1)


def test_ddpm_golden(golden):
    arrays, meta = golden
    x_T, noise = ddpm_golden_inputs()
    x = oracle.sample_latents(8, x_T, noise)
    np.testing.assert_allclose(x, arrays["ddpm_fp32"], atol=2e-5, rtol=0)
    assert meta["ddpm_fp32_vs_fp64_maxabs"] < 2e-5       # contractive sampler: 1e-4 is meaningful
    assert np.abs(x).max() <= 1.0 + 1e-6                 # t=0 step returns the clipped x0


def test_ddpm_short_run_lowp_tracks_fp32():
    x_T, noise = ddpm_golden_inputs(n=4, steps=50)
    a = oracle.sample_latents(4, x_T, noise, steps=50)
    b = oracle.sample_latents(4, x_T, noise, steps=50, lowp=torch.bfloat16)
    assert np.abs(a - b).max() < 5e-2


def test_philox_known_answers_and_normals():
    """Random123 known-answer vectors for philox4x32-10 pin the counter-based generator that the
    seeded sampler uses in-kernel (csrc/philox.cuh) and that oracle/philox.py restates."""
    def run(ctr, key):
        return [int(v) for v in oracle.philox4x32_10(np.array(ctr, dtype=np.uint32), key)]
    assert run([0, 0, 0, 0], (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run([0xFFFFFFFF] * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    z = oracle.philox_normal_rows(7, 64, 0, 8)
    assert z.shape == (8, 64, 256) and z.dtype == np.float32
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01 and np.isfinite(z).all()
    # rows are addressed by (t, latent): a sub-range reproduces the same numbers
    assert np.array_equal(oracle.philox_normal_rows(7, 64, 3, 5), z[3:5])
    assert np.array_equal(oracle.philox_normal_rows(7, 10, 3, 5, first_latent=20), z[3:5, 20:30])   # addressed by global latent index
    x_T, noise = oracle.philox_sampler_inputs(7, 64, 8)
    assert np.array_equal(noise, z) and x_T.shape == (64, 256) and not np.array_equal(x_T, z[0])


def test_synthetic_parameters_equal_the_oracle_weights(pkg):
    """bench.py's random-init parameters (product-side generator) are byte-identical to the frozen
    oracle weights, so the benchmark never has to import the oracle for its inputs."""
    assert np.array_equal(pkg.synthetic.decoder_params(), oracle.flatten_params(oracle.decoder_weights()))
    assert np.array_equal(pkg.synthetic.ddpm_params(), oracle.flatten_params(oracle.ddpm_weights()))
    for i in (0, 1, 5):
        assert np.array_equal(pkg.synthetic.latent(i), oracle.default_latent(i))


def test_latent_gradient_oracles_pin_each_other():
    """The hand-written backward of decoder_vjp_latent_lowp (what the tensor-core kernel is checked against) vs torch
    autograd on the dense forward: with lowp = float32 (no operand rounding) they agree to fp32 rounding plus the
    rare ReLU-mask flip between an fp32 and an fp64 forward (1e-3 of |grad|_max over 400 points); with bf16 / fp16
    operands the gradient of a coherent loss stays within 2 % / 0.5 % (cosine > 0.9999).  A central finite difference
    of the fp64 forward confirms the sign and size of one directional derivative; the fitting-loss oracle is the
    composition it claims to be and is invariant to the loss scale by construction (power-of-two scaling exact)."""
    import torch
    rs = np.random.RandomState(3)
    z = oracle.default_latent(2)
    xyz = (rs.rand(400, 3) * 2 - 1).astype(np.float32)
    up = np.full(400, 1.0 / 400, np.float32)
    g64, y64 = oracle.decoder_vjp_latent(z, xyz, up)
    g32, _ = oracle.decoder_vjp_latent_lowp(z, xyz, up, lowp=torch.float32)
    scale = np.abs(g64).max()
    assert np.abs(g32 - g64).max() < 1e-3 * scale
    for lowp, tol in ((torch.bfloat16, 2e-2), (torch.float16, 5e-3)):
        g, _ = oracle.decoder_vjp_latent_lowp(z, xyz, up, lowp=lowp)
        assert np.abs(g - g64).max() < tol * scale
        assert g @ g64 / (np.linalg.norm(g) * np.linalg.norm(g64)) > 0.9999
        g4, _ = oracle.decoder_vjp_latent_lowp(z, xyz, 4 * up, lowp=lowp)
        assert np.array_equal(g4, 4 * g)
    d = rs.standard_normal(256)
    d /= np.linalg.norm(d)
    eps = 1e-4
    f = lambda zz: float((oracle.decoder_forward(zz, xyz, dtype=torch.float64).astype(np.float64) * up).sum())
    fd = (f(z.astype(np.float64) + eps * d) - f(z.astype(np.float64) - eps * d)) / (2 * eps)
    assert abs(fd - float(g64 @ d)) < 1e-3 * abs(fd) + 1e-9
    tgt = oracle.decoder_forward(oracle.default_latent(7), xyz)
    loss, g = oracle.fit_loss_grad_lowp(z, xyz, tgt, clamp=0.1)
    y = oracle.decoder_forward_lowp(z, xyz)
    diff = np.clip(y, -0.1, 0.1) - np.clip(tgt, -0.1, 0.1)
    assert abs(loss - np.abs(diff).mean()) < 1e-7
    up2 = (np.sign(diff) * (np.abs(y) < 0.1) / 400).astype(np.float32)
    g2, _ = oracle.decoder_vjp_latent_lowp(z, xyz, up2)
    assert np.array_equal(g, g2)


def test_latent_gradient_golden_vectors():
    """tests/golden/vjp_golden.npz (python -m oracle.make_golden --vjp) pins the gradient oracles: recomputed here they
    match the committed vectors to CPU-BLAS summation order (fp64: 1e-9; lowp emulations: 2e-3 of |grad|_max - a
    different sgemm blocking may flip a 16-bit rounding or two)."""
    import os
    import torch
    from oracle.make_golden import GOLDEN_DIR, vjp_golden_inputs
    gold = dict(np.load(os.path.join(GOLDEN_DIR, "vjp_golden.npz")))
    xyz, up, target = vjp_golden_inputs()
    assert np.array_equal(xyz, gold["xyz"]) and np.array_equal(up, gold["up"])
    np.testing.assert_allclose(target, gold["target"], atol=2e-6, rtol=0)
    z = oracle.default_latent(2)
    g64, y64 = oracle.decoder_vjp_latent(z, xyz, up)
    np.testing.assert_allclose(g64, gold["grad_fp64"], atol=1e-9 * np.abs(gold["grad_fp64"]).max(), rtol=0)
    for name, lowp in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        g, y = oracle.decoder_vjp_latent_lowp(z, xyz, up, lowp=lowp)
        scale = np.abs(gold[f"grad_{name}"]).max()
        assert np.abs(g - gold[f"grad_{name}"]).max() < 2e-3 * scale
        loss, gf = oracle.fit_loss_grad_lowp(z, xyz, gold["target"], clamp=0.1, lowp=lowp)
        assert abs(loss - float(gold[f"fit_loss_{name}"])) < 2e-5
        assert np.abs(gf - gold[f"fit_grad_{name}"]).max() < 2e-2 * np.abs(gold[f"fit_grad_{name}"]).max()


def test_training_step_golden_vectors():
    """tests/golden/train_golden.npz (python -m oracle.make_golden --train) pins the training-step oracles of SURVEY 8f row N4
    (oracle/train.py): recomputed here, loss, per-tensor gradient norms and 4096 sampled gradient entries match the committed
    vectors to CPU-BLAS summation order (fp64: 1e-9 relative; the emulations of the 16-bit step: a flipped rounding or ReLU
    mask may move single entries by a percent of |grad|_max, the norms agree to 1e-3)."""
    import os
    import torch
    from oracle.make_golden import GOLDEN_DIR, train_golden_inputs, train_sample_indices
    gold = dict(np.load(os.path.join(GOLDEN_DIR, "train_golden.npz")))
    (x0, t, eps), (lat, xyz, tgt) = train_golden_inputs()

    def check(prefix, loss, grads, tol_entry, tol_norm):
        flat = oracle.flatten_grads(grads)
        smp = flat[train_sample_indices(flat.size)]
        ref = gold[prefix + "_grad_sample"]
        assert abs(loss - float(gold[prefix + "_loss"])) <= tol_norm * max(1.0, abs(float(gold[prefix + "_loss"])))
        assert np.abs(smp - ref).max() <= tol_entry * np.abs(ref).max(), prefix
        norms = np.array([np.linalg.norm(np.asarray(g, dtype=np.float64)) for pair in grads for g in pair])
        np.testing.assert_allclose(norms, gold[prefix + "_tensor_norms"], rtol=tol_norm, atol=tol_norm * float(gold[prefix + "_grad_norm"]))

    l, g = oracle.ddpm_train_grads(x0, t, eps)
    check("ddpm_fp64", l, g, 1e-6, 1e-6)           # (the stored samples are fp32)
    l, g, _ = oracle.decoder_train_grads(lat, xyz, tgt)
    check("dec_fp64", l, g, 1e-6, 1e-6)
    for name, lowp in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        l, g = oracle.ddpm_train_grads_lowp(x0, t, eps, lowp=lowp)
        check("ddpm_" + name, l, g, 5e-2, 2e-3)
        l, g, y = oracle.decoder_train_grads_lowp(lat, xyz, tgt, lowp=lowp)
        check("dec_" + name, l, g, 5e-2, 2e-3)
        assert np.abs(np.asarray(y, np.float32).ravel() - gold["dec_" + name + "_sdf"]).max() < 2e-3
