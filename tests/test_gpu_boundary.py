"""Boundary behaviour of the C ABI on a GPU: watchdog trips surface on the next call, host-buffer calls are ordered after
device-path calls on the same context, empty slabs are accepted, the host mask is the device mask, and a plain C program
decodes a grid through `sdfb_decode_grid_host` and matches a committed golden field.  (No reference interface exists to
mirror - /root/reference/README.md:1; the contract is include/sdfb200.h.)"""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params():
    return oracle.flatten_params(oracle.decoder_weights())


def test_watchdog_trip_surfaces_on_the_next_call(pkg, monkeypatch):
    """SDFB_DEBUG_FLAGS=2: the weight producer exits at once, so every consumer wait runs into the watchdog.  The async
    call itself returns 0 (its kernel has not run yet); the NEXT call on the context returns SDFB_E_KERNEL, once."""
    z = torch.from_numpy(oracle.default_latent()).cuda()
    good = pkg.Decoder(_params(), device="cuda:0")
    ref = good.decode_grid(z, 32)
    monkeypatch.setenv("SDFB_DEBUG_FLAGS", "2")
    bad = pkg.Decoder(_params(), device="cuda:0")
    monkeypatch.delenv("SDFB_DEBUG_FLAGS")
    bad.set_watchdog_timeout_ns(2_000_000)                 # 2 ms per wait
    bad.decode_grid(z, 32)                                  # asynchronous: no error yet
    torch.cuda.synchronize()
    with pytest.raises(pkg.SdfbError) as e:
        bad.decode_grid(z, 32)
    assert e.value.code == -4 and "watchdog" in str(e.value)
    bad.decode_grid(z, 32)                                  # reported once, then cleared; this launch fails again ...
    with pytest.raises(pkg.SdfbError):
        bad.check()                                         # ... and check() (synchronising) reports it
    bad.close()
    good.check()                                            # an unrelated context is unaffected
    assert torch.equal(good.decode_grid(z, 32), ref)


def test_host_call_is_ordered_after_device_calls_on_the_same_context(cuda_decoder):
    """decode_grid(zA) (async, caller's stream) followed by decode_grid_host(zB) (context's own streams): the second call
    rewrites the folded latent constants and must not do so while the first kernel is still reading them."""
    zA, zB = oracle.default_latent(0), oracle.default_latent(5)
    refA = cuda_decoder.decode_grid(zA, 160).clone()
    refB = cuda_decoder.decode_grid(zB, 64).cpu().numpy()
    torch.cuda.synchronize()
    for _ in range(3):
        a = cuda_decoder.decode_grid(zA, 160)               # ~9 ms of kernel, still running when the host call starts
        b = cuda_decoder.decode_grid_host(zB, 64)
        assert np.array_equal(b, refB)
        assert torch.equal(a, refA)


def test_empty_slab_is_accepted(cuda_decoder, pkg):
    """slab_range hands tail ranks an empty range for some sizes (res 9 over 4 ranks: rank 3 gets [9, 9))."""
    z = oracle.default_latent()
    assert pkg.slab_range(9, 3, 4) == (9, 9)
    sdf, m = cuda_decoder.decode_grid(z, 9, 9, 9, mask=True)
    assert sdf.shape == (0, 9, 9) and m.shape == (0, 8, 8)
    assert cuda_decoder.decode_grid(z, 9, 9, 9).shape == (0, 9, 9)
    whole = cuda_decoder.decode_grid(z, 9)
    parts = [cuda_decoder.decode_grid(z, 9, *pkg.slab_range(9, r, 4)) for r in range(4)]
    assert torch.equal(torch.cat(parts), whole)
    lib = pkg.load_library()
    zt = torch.from_numpy(z).cuda()
    assert lib.sdfb_decode_grid(cuda_decoder._h, zt.data_ptr(), 9, 9, 9, None, None, 1, None) == 0
    assert lib.sdfb_decode_grid(cuda_decoder._h, zt.data_ptr(), 9, 3, 9, None, None, 1, None) == -1


@pytest.mark.parametrize("res,z0,z1", [(64, 0, 64), (96, 10, 50), (33, 0, 33), (33, 5, 20)])
def test_host_mask_equals_device_mask(cuda_decoder, res, z0, z1):
    """The host-buffer call builds its mask from the sign bit-planes the chunked launches write (no second pass over the
    fp32 field): same bits as the device path and as the oracle's mask function of the same field."""
    z = oracle.default_latent(2)
    sdf_d, m_d = cuda_decoder.decode_grid(z, res, z0, z1, mask=True)
    sdf_h, m_h = cuda_decoder.decode_grid_host(z, res, z0, z1, mask=True)
    assert np.array_equal(sdf_h, sdf_d.cpu().numpy())
    assert np.array_equal(m_h, m_d.cpu().numpy())
    if z1 == res:
        assert np.array_equal(m_h, oracle.sign_change_mask(sdf_h))


def test_c_program_decodes_through_the_host_entry_point(tmp_path):
    """tests/c/gpu_decode.c (plain C99): create a decoder from a parameter blob, sdfb_decode_grid_host a 32^3 grid in fp32
    and bf16, compare with the committed golden field (tests/golden/c_abi_decode32.bin, oracle fp32)."""
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    lib = os.path.join(ROOT, "latent-diffusion-models-for-shape-sdfs_b200", "libsdfb200.so")
    exe = str(tmp_path / "gpu_decode")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "gpu_decode.c"), "-o", exe, lib, "-lm", "-Wl,-rpath," + os.path.dirname(lib)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    blob = str(tmp_path / "params.bin")
    _params().astype(np.float32).tofile(blob)
    lat = str(tmp_path / "latent.bin")
    oracle.default_latent().astype(np.float32).tofile(lat)
    run = subprocess.run([exe, blob, lat, os.path.join(ROOT, "tests", "golden", "c_abi_decode32.bin")],
                         capture_output=True, text=True, timeout=300)
    print(run.stdout)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "gpu_decode ok" in run.stdout


def test_round2_entry_points_reject_bad_arguments(pkg):
    """Error behaviour of the entry points added in round 2: SDFB_E_INVALID (< 0) and a message, never a launch."""
    import ctypes as C
    lib = pkg.load_library()
    z = torch.zeros((2, 256), device="cuda")
    n = C.c_int64()
    nbytes = C.c_size_t()
    # welding: null workspace, grid too small, negative triangle count, workspace too small
    assert lib.sdfb_mc_weld_workspace_bytes(1, C.byref(nbytes)) < 0
    assert lib.sdfb_mc_weld_workspace_bytes(16, C.byref(nbytes)) == 0 and nbytes.value > 0
    ws = torch.empty((nbytes.value,), dtype=torch.uint8, device="cuda")
    assert lib.sdfb_mc_weld_count(None, 0, 16, None, nbytes.value, C.byref(n), None) < 0
    assert lib.sdfb_mc_weld_count(None, -1, 16, ws.data_ptr(), nbytes.value, C.byref(n), None) < 0
    assert lib.sdfb_mc_weld_count(None, 0, 16, ws.data_ptr(), 8, C.byref(n), None) < 0
    assert lib.sdfb_mc_weld_count(None, 0, 16, ws.data_ptr(), nbytes.value, C.byref(n), None) == 0 and n.value == 0      # empty soup
    assert lib.sdfb_mc_weld_fill(None, None, 0, 16, ws.data_ptr(), None, None, None) == 0
    assert lib.sdfb_mc_weld_fill(None, None, 3, 16, ws.data_ptr(), None, None, None) < 0
    # latent Adam: step 0, betas outside [0, 1), null moments
    assert lib.sdfb_latent_adam_step(z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), None, 2, 1e-3, 0.0, 0.9, 0.999, 1e-8, 0, None) < 0
    assert lib.sdfb_latent_adam_step(z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), None, 2, 1e-3, 0.0, 1.0, 0.999, 1e-8, 1, None) < 0
    assert lib.sdfb_latent_adam_step(z.data_ptr(), None, z.data_ptr(), z.data_ptr(), None, 2, 1e-3, 0.0, 0.9, 0.999, 1e-8, 1, None) < 0
    # diagnostics: the probes refuse shapes their protocols cannot run
    out3 = (C.c_double * 3)()
    assert lib.sdfb_tma_ingest_rate(148, 3, 0, 1, 0, 64, 32, 64, out3) < 0           # cluster size not a power of two
    assert lib.sdfb_tma_ingest_rate(148, 1, 9, 1, 0, 64, 32, 64, out3) < 0           # unknown mode
    assert lib.sdfb_tma_ingest_rate(128, 4, 5, 1, 0, 64, 32, 64, out3) < 0           # pair modes need clusters of 2
    v = C.c_double()
    assert lib.sdfb_umma_rate(1, 2, 10, 4, 2, 8, C.byref(v)) < 0                     # the N = 128 form with intermediate waits
    assert b"N = 128" in lib.sdfb_last_error()
    torch.cuda.synchronize()


def test_integration_md_binding_runs_as_written(pkg, cuda_decoder):
    """The reference-side ctypes binding shown in INTEGRATION.md ("Minimal binding"), executed verbatim from the repo root
    with the oracle's seeded weights: same results as the package's own wrapper, and the samplers run."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"## Minimal binding.*?```python\n(.*?)```", text, re.S).group(1)
    ns = {"decoder_layers": oracle.decoder_weights(), "denoiser_layers": oracle.ddpm_weights()}
    cwd = os.getcwd()
    os.chdir(root)
    try:
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
        z = oracle.default_latent()
        sdf, mask = ns["decode_grid"](z, 32)
        ref_sdf, ref_mask = cuda_decoder.decode_grid(z, 32, mask=True, precision="bf16")
        assert np.array_equal(sdf, ref_sdf.cpu().numpy()) and np.array_equal(mask, ref_mask.cpu().numpy().reshape(mask.shape))
        pts = (np.random.RandomState(0).rand(100, 3) * 2 - 1).astype(np.float32)
        y = ns["Decoder"](z, pts, precision=0)
        assert np.abs(y - oracle.decoder_forward(z, pts)).max() < 1e-5
        x = ns["sample_latents"](8, seed=3, steps=6)
        assert x.shape == (8, 256) and np.isfinite(x).all() and np.abs(x).max() <= 1.0 + 1e-6
        x_T, noise = np.zeros((4, 256), np.float32), np.zeros((3, 4, 256), np.float32)
        x2 = ns["sample_latents_explicit"](x_T, noise, precision=0)
        assert np.abs(x2 - oracle.sample_latents(4, x_T, noise, steps=3)).max() < 1e-4
    finally:
        os.chdir(cwd)


def test_contexts_release_their_device_memory(pkg):
    """Create / use / destroy every kind of context a few times: free device memory comes back (no leak per context)."""
    import gc
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    def round_trip():
        dec = pkg.Decoder(pkg.synthetic.decoder_params(), device="cuda:0", precision="bf16")
        z = torch.from_numpy(pkg.synthetic.latent(0)).cuda()
        dec.decode_grid(z, 48, mask=True)
        dec.extract_surface_sparse(z, 65)
        dd = pkg.LatentDDPM(pkg.synthetic.ddpm_params(), device="cuda:0", precision="bf16")
        dd.sample_latents(300, steps=3, seed=1)
        tr = pkg.DDPMTrainer(pkg.synthetic.ddpm_params(), precision="bf16")
        x0 = torch.zeros((64, 256), device="cuda")
        tr.step(x0, torch.zeros(64, dtype=torch.int32, device="cuda"), x0)
        dt = pkg.DecoderTrainer(pkg.synthetic.decoder_params(), precision="bf16")
        dt.step(z[None], torch.zeros((1, 64, 3), device="cuda"), torch.zeros((1, 64), device="cuda"), lr=1e-6)
        torch.cuda.synchronize()
        for c in (dec, dd, tr, dt):
            c.close()
        del dec, dd, tr, dt, z, x0
        gc.collect()
        torch.cuda.empty_cache()

    round_trip()                                    # first use: lazily created library state stays
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(4):
        round_trip()
    free1, _ = torch.cuda.mem_get_info()
    print(f"free device memory before / after 4 context round trips: {free0 >> 20} / {free1 >> 20} MiB")
    assert free0 - free1 < (64 << 20)
