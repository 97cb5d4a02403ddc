"""SURVEY.md section 8f row N4 (first step): the vector-Jacobian product of the decoder w.r.t. the latent
on the fp32 path, against torch autograd on the oracle's dense forward, and latent fitting built on it."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def test_latent_vjp_matches_autograd(cuda_decoder):
    """Tolerances.  The gradient of a ReLU network is piecewise constant in its masks: an fp32 and an fp64
    forward disagree on a few unit masks per 10^7 activations (|pre-activation| ~ 1e-7), and each flip moves the
    summed gradient by ~1e-3.  Small batches (no flips): 2e-5 relative.  Large batches: the CUDA path must be as
    close to the fp64 autograd as torch's own fp32 autograd is (same order of magnitude, cosine > 0.99999)."""
    rs = np.random.RandomState(11)
    z = oracle.default_latent(2)
    for M in (1, 300, 9000):                      # 9000 > one 8192-row chunk
        xyz = (rs.rand(M, 3) * 2 - 1).astype(np.float32)
        up = rs.standard_normal(M).astype(np.float32)
        g, y = cuda_decoder.latent_vjp(z, xyz, up)
        g = g.cpu().numpy().astype(np.float64)
        g_ref, y_ref = oracle.decoder_vjp_latent(z, xyz, up)              # float64 autograd
        scale = np.abs(g_ref).max()
        err = np.abs(g - g_ref).max()
        cos = float(g @ g_ref / (np.linalg.norm(g) * np.linalg.norm(g_ref)))
        assert np.abs(y.cpu().numpy() - y_ref).max() < 1e-5
        if M <= 300:
            print(f"M={M}: max|grad - autograd(fp64)| = {err:.3e} (|grad|_max {scale:.3e})")
            assert err < 2e-5 * max(scale, 1.0)
        else:
            g32, _ = oracle.decoder_vjp_latent(z, xyz, up, dtype=torch.float32)
            err32 = np.abs(g32 - g_ref).max()
            print(f"M={M}: max|grad - autograd(fp64)| = {err:.3e}, torch fp32 autograd vs fp64 = {err32:.3e} "
                  f"(|grad|_max {scale:.3e}), cosine {cos:.8f}")
            assert err < max(5 * err32, 5e-3 * scale) and cos > 0.99999
            # chunking: the sum of two separate calls equals the single call up to fp32 reduction order
            ga, _ = cuda_decoder.latent_vjp(z, xyz[:8192], up[:8192])
            gb, _ = cuda_decoder.latent_vjp(z, xyz[8192:], up[8192:])
            assert float((ga + gb).cpu().sub(torch.from_numpy(g).float()).abs().max()) < 2e-4
    g0, _ = cuda_decoder.latent_vjp(z, xyz[:0], up[:0])
    assert float(g0.abs().max()) == 0.0
    g1, _ = cuda_decoder.latent_vjp(z, xyz, up)
    g2, _ = cuda_decoder.latent_vjp(z, xyz, up)
    assert torch.equal(g1, g2)                    # deterministic reduction


def test_fit_latent_recovers_a_shape(cuda_decoder):
    """Auto-decoder inference: fit a latent to SDF samples drawn from a known shape's field."""
    rs = np.random.RandomState(5)
    z_true = oracle.default_latent(7)
    xyz = (rs.rand(20000, 3) * 2 - 1).astype(np.float32)
    tgt = cuda_decoder(z_true, xyz, precision="fp32")
    base = float((torch.clamp(cuda_decoder(np.zeros(256, np.float32), xyz, precision="fp32"), -0.1, 0.1)
                  - torch.clamp(tgt, -0.1, 0.1)).abs().mean())
    z_fit, loss = cuda_decoder.fit_latent(xyz, tgt, steps=200, lr=1e-2, reg=0.0)
    test_xyz = (rs.rand(5000, 3) * 2 - 1).astype(np.float32)
    err = float((torch.clamp(cuda_decoder(z_fit, test_xyz, precision="fp32"), -0.1, 0.1)
                 - torch.clamp(cuda_decoder(z_true, test_xyz, precision="fp32"), -0.1, 0.1)).abs().mean())
    frac_in = float((tgt < 0).float().mean())
    print(f"fit_latent: clamped-L1 {base:.5f} (zero latent) -> {loss:.5f} (train) / {err:.5f} (held-out points); "
          f"target shape has {frac_in:.2f} of the samples inside")
    assert 0.05 < frac_in < 0.6              # a real surface, not a saturated field
    assert err < 0.25 * base
