"""GPU parity tests of the latent DDPM sampler (C ABI -> CUDA) against the oracle under an
identical, explicit noise stream.  Tolerance: 1e-4 on the fp32 path (north_star)."""
import numpy as np
import pytest
import torch

import oracle
from oracle.make_golden import ddpm_golden_inputs

pytestmark = pytest.mark.gpu


def test_denoiser_single_step_fp32(cuda_ddpm):
    rs = np.random.RandomState(3)
    x = rs.standard_normal((37, 256)).astype(np.float32)
    for t in (0, 1, 500, 999):
        eps = cuda_ddpm.denoise(x, t, precision="fp32").cpu().numpy()
        ref = oracle.denoiser_forward(x, t).numpy()
        assert np.abs(eps - ref).max() < 2e-5, (t, np.abs(eps - ref).max())


def test_sample_latents_golden_fp32(cuda_ddpm, golden):
    arrays, _ = golden
    x_T, noise = ddpm_golden_inputs()
    x = cuda_ddpm.sample_latents(8, x_T=x_T, noise=noise, precision="fp32").cpu().numpy()
    err = np.abs(x - arrays["ddpm_fp32"]).max()
    print(f"ddpm fp32 1000 steps: max|kernel - oracle| = {err:.3e}")
    assert err < 1e-4
    assert np.abs(x).max() <= 1.0 + 1e-6
    xh = cuda_ddpm.sample_latents_host(x_T, noise, precision="fp32")
    assert np.array_equal(xh, x)


def test_sample_latents_short_runs_and_device_noise(cuda_ddpm):
    x_T, noise = ddpm_golden_inputs(n=5, steps=30)
    x = cuda_ddpm.sample_latents(5, x_T=x_T, noise=noise, steps=30, precision="fp32").cpu().numpy()
    ref = oracle.sample_latents(5, x_T, noise, steps=30)
    assert np.abs(x - ref).max() < 1e-4
    a = cuda_ddpm.sample_latents(3, steps=10, seed=9)
    b = cuda_ddpm.sample_latents(3, steps=10, seed=9)
    assert torch.equal(a, b) and a.shape == (3, 256)
    with pytest.raises(ValueError):
        cuda_ddpm.sample_latents(0)
