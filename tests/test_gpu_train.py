"""GPU parity tests of the training steps (second half) through the C ABI: the general
tensor-core product (both operand layouts), the DDPM denoiser training step against the fp64-autograd oracle and against
the oracle that emulates the step's 16-bit roundings (oracle/train.py), Adam against its numpy restatement.
Tolerances: the emulating oracle and the kernel differ only in fp32 summation order (and the rounding flips it causes), so
gradients agree in direction to ~1e-5 (single entries move by a few percent of the largest where a ReLU mask flips); against fp64 autograd the 16-bit operands cost a cosine of ~1e-3 (bf16) /
~1e-6 (fp16).  No upstream source exists (/root/reference/README.md:1)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
LOWP_T = {"bf16": torch.bfloat16, "fp16": torch.float16}


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("tn,M,N,K,bn,ksplit", [(0, 128, 256, 64, 256, 1), (0, 300, 1024, 512, 256, 1), (0, 77, 256, 1024, 128, 1),
                                                 (0, 4096, 64, 320, 64, 1), (1, 128, 256, 64, 256, 1), (1, 1024, 512, 300, 256, 3),
                                                 (1, 256, 1024, 4096, 256, 16), (1, 512, 320, 1000, 64, 4)])
def test_general_product_both_layouts(pkg, prec, tn, M, N, K, bn, ksplit):
    """out = a . b^T (K-major operands) and out = a^T . b (MN-major operands: the contraction runs over the ROWS of two
    row-major arrays, which is what a weight gradient is) against torch in fp32; ragged M and K, split K."""
    lib = pkg.load_library()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K + tn)
    dt = LOWP_T[prec]
    if tn:
        a = torch.randn((K, M), generator=g, device="cuda").to(dt)
        b = torch.randn((K, N), generator=g, device="cuda").to(dt)
        ref = a.float().T @ b.float()
        lda, ldb = M, N
    else:
        a = torch.randn((M, K), generator=g, device="cuda").to(dt)
        b = torch.randn((N, K), generator=g, device="cuda").to(dt)
        ref = a.float() @ b.float().T
        lda, ldb = K, K
    out = torch.full((ksplit, M, N), float("nan"), device="cuda")
    rc = lib.sdfb_gemm_selftest(a.data_ptr(), lda, b.data_ptr(), ldb, M, N, K, tn, ksplit, bn, pkg.PRECISIONS[prec], out.data_ptr(), None)
    assert rc == 0, lib.sdfb_last_error()
    got = out.sum(dim=0)
    err = (got - ref).abs().max().item()
    print(f"{prec} tn={tn} {M}x{N}x{K} bn={bn} ksplit={ksplit}: max err {err:.2e} (|ref| max {ref.abs().max().item():.1f})")
    assert err < 2e-3 * max(1.0, ref.abs().max().item() / 50), err


def _batch(n, seed):
    rs = np.random.RandomState(seed)
    x0 = np.clip(rs.standard_normal((n, 256)) * 0.6, -1, 1).astype(np.float32)
    eps = rs.standard_normal((n, 256)).astype(np.float32)
    t = rs.randint(0, 1000, n).astype(np.int32)
    return x0, t, eps


@pytest.mark.parametrize("prec,n", [("bf16", 64), ("fp16", 64), ("bf16", 700), ("fp16", 1)])
def test_ddpm_training_step_gradients(pkg, prec, n):
    params = oracle.flatten_params(oracle.ddpm_weights())
    tr = pkg.DDPMTrainer(params, precision=prec)
    x0, t, eps = _batch(n, 11 + n)
    loss, grads = tr.step(x0, t, eps, apply=False, return_grads=True)
    loss, g = float(loss.item()), grads.cpu().numpy()
    l_low, g_low = oracle.ddpm_train_grads_lowp(x0, t, eps, lowp=LOWP_T[prec])
    l_64, g_64 = oracle.ddpm_train_grads(x0, t, eps)
    g_low, g_64 = oracle.flatten_grads(g_low), oracle.flatten_grads(g_64).astype(np.float64)
    gmax = np.abs(g_64).max()
    e_low = np.abs(g - g_low).max() / gmax
    cos64 = float(g.astype(np.float64) @ g_64 / np.linalg.norm(g) / np.linalg.norm(g_64))
    cos_low = float(g.astype(np.float64) @ g_low / np.linalg.norm(g) / np.linalg.norm(g_low))
    print(f"{prec} n={n}: loss {loss:.6f} (emulating oracle {l_low:.6f}, fp64 {l_64:.6f}); max|grad - emulating oracle| = {e_low:.2e} of |grad|_max, "
          f"cosine {cos_low:.7f}; vs fp64 autograd cosine {cos64:.7f}")
    assert abs(loss - l_low) < 2e-4 * max(1.0, l_low) and abs(loss - l_64) < 2e-2 * max(1.0, l_64)
    # single entries move by a few percent of |grad|_max where a summation-order difference flips a ReLU mask (the same
    # effect as in the latent-gradient tests); the direction is what training uses
    assert e_low < 8e-2 and cos_low > 0.9998
    assert cos64 > (0.9995 if prec == "fp16" else 0.995)
    assert np.array_equal(tr.params(), params)                       # apply=False leaves the parameters alone
    tr.close()


def test_adam_update_and_a_short_training_run(pkg):
    """apply=True: parameters move by exactly the Adam update of the gradient the step reports (numpy restatement), the 16-bit
    weight copies follow (the next step's loss reflects the update), and a few steps on a fixed batch bring the loss down."""
    params = oracle.flatten_params(oracle.ddpm_weights())
    tr = pkg.DDPMTrainer(params, precision="bf16")
    x0, t, eps = _batch(512, 3)
    m = np.zeros_like(params)
    v = np.zeros_like(params)
    p = params.copy()
    losses = []
    for step in range(1, 4):
        loss, grads = tr.step(x0, t, eps, lr=1e-3, apply=True, return_grads=True)
        losses.append(float(loss.item()))
        p, m, v = oracle.adam_step(p, grads.cpu().numpy(), m, v, step, 1e-3)
        got = tr.params()
        err = np.abs(got - p).max()
        print(f"step {step}: loss {losses[-1]:.5f}, max|params - numpy Adam| = {err:.2e}")
        assert err < 2e-6
    for _ in range(40):
        loss = tr.step(x0, t, eps, lr=1e-3)
    losses.append(float(loss.item()))
    print("loss after 43 steps on a fixed batch:", losses[-1])
    assert losses[-1] < 0.7 * losses[0]
    smp = tr.sampler(precision="bf16")                                # the trained weights drive the fused sampler
    x = smp.sample_latents(64, steps=20, seed=5)
    assert torch.isfinite(x).all()
    smp.close()
    tr.close()


def _shapes(B, P, seed):
    rs = np.random.RandomState(seed)
    lat = np.stack([oracle.default_latent(i) for i in range(B)])
    xyz = (rs.rand(B, P, 3) * 2 - 1).astype(np.float32)
    tgt = np.stack([oracle.decoder_forward(oracle.default_latent(7 + i), xyz[i]) for i in range(B)])
    return lat, xyz, tgt


@pytest.mark.parametrize("prec,B,P", [("bf16", 1, 64), ("fp16", 3, 200), ("bf16", 4, 1000), ("fp16", 2, 37)])
def test_decoder_weight_gradients(pkg, prec, B, P):
    """dW_l = delta_l^T a_l for all nine layers (the head in fp32), a batch of shapes with their own latents: against the oracle
    that emulates the step's roundings and against fp64 autograd (oracle/train.py decoder_train_grads*)."""
    params = oracle.flatten_params(oracle.decoder_weights())
    tr = pkg.DecoderTrainer(params, precision=prec)
    lat, xyz, tgt = _shapes(B, P, 5 + P)
    loss, grads, sdf = tr.step(lat, xyz, tgt, apply=False, return_grads=True, return_sdf=True)
    loss, g, sdf = float(loss.item()), grads.cpu().numpy(), sdf.cpu().numpy().ravel()
    l_low, g_low, y_low = oracle.decoder_train_grads_lowp(lat, xyz, tgt, lowp=LOWP_T[prec])
    l_64, g_64, y_64 = oracle.decoder_train_grads(lat, xyz, tgt)
    g_low, g_64 = oracle.flatten_grads(g_low), oracle.flatten_grads(g_64).astype(np.float64)
    gmax = np.abs(g_64).max()
    e_low = np.abs(g - g_low).max() / gmax
    cos64 = float(g.astype(np.float64) @ g_64 / np.linalg.norm(g) / np.linalg.norm(g_64))
    cos_low = float(g.astype(np.float64) @ g_low / np.linalg.norm(g) / np.linalg.norm(g_low))
    print(f"{prec} B={B} P={P}: loss {loss:.6f} (emulating oracle {l_low:.6f}, fp64 {l_64:.6f}); max|sdf - emulating oracle| {np.abs(sdf - y_low).max():.2e}; "
          f"max|grad - emulating oracle| = {e_low:.2e} of |grad|_max, cosine {cos_low:.7f}; vs fp64 autograd cosine {cos64:.7f}")
    assert np.abs(sdf - y_low).max() < (8e-3 if prec == "bf16" else 1.5e-3)
    assert abs(loss - l_low) < 5e-4 and abs(loss - l_64) < 5e-3
    assert e_low < 8e-2 and cos_low > 0.999
    if B * P >= 500:                 # (with a handful of samples single ReLU-mask flips of the 16-bit forward dominate the direction)
        assert cos64 > (0.9995 if prec == "fp16" else 0.999)
    assert np.array_equal(tr.params(), params)
    tr.close()


def test_decoder_training_run(pkg):
    """Adam on the decoder's weights: parameters follow the numpy restatement of Adam applied to the reported gradient, and a
    few steps on a fixed batch of (shape, samples) pairs bring the clamped-L1 loss down; the trained weights load into the
    fused inference kernel."""
    params = oracle.flatten_params(oracle.decoder_weights())
    tr = pkg.DecoderTrainer(params, precision="bf16")
    lat, xyz, tgt = _shapes(4, 4096, 21)
    m, v, p = np.zeros_like(params), np.zeros_like(params), params.copy()
    losses = []
    for step in range(1, 3):
        loss, grads = tr.step(lat, xyz, tgt, lr=1e-5, apply=True, return_grads=True)
        losses.append(float(loss.item()))
        p, m, v = oracle.adam_step(p, grads.cpu().numpy(), m, v, step, 1e-5)
        err = np.abs(tr.params() - p).max()
        print(f"step {step}: loss {losses[-1]:.5f}, max|params - numpy Adam| = {err:.2e}")
        assert err < 2e-6
    for _ in range(60):             # (lr 1e-5: Adam moves every weight by ~lr per step whatever the gradient's size, and this
        loss = tr.step(lat, xyz, tgt, lr=1e-5)   # random-init network is pushed out of the clamp band by steps of 2e-4)
    losses.append(float(loss.item()))
    print("clamped-L1 after 62 steps on a fixed batch:", losses[-1], "from", losses[0])
    assert losses[-1] < 0.6 * losses[0]
    dec = tr.decoder()
    y = dec(lat[0], xyz[0]).cpu().numpy()
    assert np.isfinite(y).all()
    l0 = np.abs(np.clip(y, -0.1, 0.1) - np.clip(tgt[0], -0.1, 0.1)).mean()
    print("shape 0 clamped-L1 through the fused inference kernel with the trained weights:", l0)
    assert l0 < 0.8 * losses[0]
    dec.close()
    tr.close()


def test_latent_adam_step_is_the_tensor_expression(pkg):
    """sdfb_latent_adam_step (one launch per fitting step) against the fp32 tensor expression it replaced in
    Decoder.fit_latent / fit_latents_batch: the same moments bit for bit over several steps, the latents to an ulp or two
    per step (the library's elementwise division is not the correctly rounded one), the reported loss (a 256-term sum in
    another order) to rounding."""
    lib = pkg.load_library()
    g0 = torch.Generator(device="cuda").manual_seed(3)
    B, lr, reg = 5, 5e-3, 1e-4
    z = 0.1 * torch.randn((B, 256), generator=g0, device="cuda")
    m, v = torch.zeros_like(z), torch.zeros_like(z)
    z2, m2, v2 = z.clone(), m.clone(), v.clone()
    for it in range(1, 8):
        g = torch.randn((B, 256), generator=g0, device="cuda") * (10.0 ** -(it % 4))
        loss0 = torch.rand(B, generator=g0, device="cuda")
        # reference: api.py of round 1
        loss_ref = loss0 + reg * (z * z).sum(dim=1)
        gg = g + 2 * reg * z
        m = 0.9 * m + 0.1 * gg
        v = 0.999 * v + 0.001 * gg * gg
        z = z - lr * (m / (1 - 0.9 ** it)) / ((v / (1 - 0.999 ** it)).sqrt() + 1e-8)
        loss = loss0.clone()
        rc = lib.sdfb_latent_adam_step(z2.data_ptr(), m2.data_ptr(), v2.data_ptr(), g.data_ptr(), loss.data_ptr(), B, lr, reg, 0.9, 0.999,
                                       1e-8, it, None)
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(m2, m), f"first moments differ at step {it}: {(m2 - m).abs().max().item():.3e}"
        assert torch.equal(v2, v), f"second moments differ at step {it}: {(v2 - v).abs().max().item():.3e} of {v.abs().max().item():.3e}"
        assert torch.allclose(z2, z, rtol=0, atol=it * 1e-7), f"latents differ at step {it}: {(z2 - z).abs().max().item():.3e}"
        z2.copy_(z)                       # keep the two trajectories on the same inputs
        assert torch.allclose(loss, loss_ref, rtol=1e-6, atol=1e-9)
    assert lib.sdfb_latent_adam_step(z2.data_ptr(), m2.data_ptr(), v2.data_ptr(), g.data_ptr(), None, B, lr, reg, 0.9, 0.999, 1e-8, 0, None) != 0
    assert lib.sdfb_latent_adam_step(None, None, None, None, None, 0, lr, reg, 0.9, 0.999, 1e-8, 1, None) == 0       # empty batch


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_training_steps_against_committed_golden_vectors(pkg, prec):
    """Both training steps against tests/golden/train_golden.npz (python -m oracle.make_golden --train; the GPU box has no
    /root/reference and need not recompute the oracle): loss, 4096 sampled gradient entries and every tensor's gradient norm."""
    import os
    from oracle.make_golden import train_golden_inputs, train_sample_indices
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz")))
    (x0, t, eps), (lat, xyz, tgt) = train_golden_inputs()

    def compare(prefix, loss, g, sizes):
        smp, ref = g[train_sample_indices(g.size)], gold[prefix + "_grad_sample"]
        err = np.abs(smp - ref).max() / np.abs(ref).max()
        cos = float(smp.astype(np.float64) @ ref / np.linalg.norm(smp) / np.linalg.norm(ref))
        norms, o = [], 0
        for n in sizes:
            norms.append(np.linalg.norm(g[o:o + n].astype(np.float64)))
            o += n
        nerr = np.abs(np.array(norms) - gold[prefix + "_tensor_norms"]).max() / float(gold[prefix + "_grad_norm"])
        print(f"{prefix}: loss {loss:.6f} vs {float(gold[prefix + '_loss']):.6f}; sampled entries max err {err:.2e} of max, cosine {cos:.7f}; "
              f"tensor norms within {nerr:.2e} of |grad|")
        assert abs(loss - float(gold[prefix + "_loss"])) < 5e-4 * max(1.0, float(gold[prefix + "_loss"]))
        assert err < 8e-2 and cos > 0.999 and nerr < 1e-2

    dp = oracle.ddpm_weights()
    tr = pkg.DDPMTrainer(oracle.flatten_params(dp), precision=prec)
    loss, grads = tr.step(x0, t, eps, apply=False, return_grads=True)
    compare("ddpm_" + prec, float(loss.item()), grads.cpu().numpy(), [np.asarray(a).size for pair in dp for a in pair])
    tr.close()
    wp = oracle.decoder_weights()
    dt = pkg.DecoderTrainer(oracle.flatten_params(wp), precision=prec)
    loss, grads, sdf = dt.step(lat, xyz, tgt, apply=False, return_grads=True, return_sdf=True)
    compare("dec_" + prec, float(loss.item()), grads.cpu().numpy(), [np.asarray(a).size for pair in wp for a in pair])
    assert np.abs(sdf.cpu().numpy().ravel() - gold["dec_" + prec + "_sdf"]).max() < (8e-3 if prec == "bf16" else 1.5e-3)
    dt.close()
