"""CPU oracle for the shape-SDF decoder / latent-DDPM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or the timed CPU baseline), never as a fallback.

Parity status: **parity unpinned by the reference.**  The mounted reference
(`/root/reference/README.md:1`) is a one-line title with no source, tests or
golden vectors, so there is nothing upstream to pin this oracle against.  The
oracle restates the method named by BASELINE.json's ``north_star`` (SURVEY.md
section 8a, rows A1-A7) and is pinned instead by (a) an fp64 re-evaluation of
itself, (b) a SHA-256 of the frozen weights and (c) the fixtures in
``tests/golden/`` produced by ``oracle/make_golden.py``.
"""
from .weights import (DEC_LATENT, DEC_HIDDEN, DEC_SKIP_OUT, DEC_LAYER_DIMS,
                      DDPM_LAYER_DIMS, DDPM_T, decoder_weights, ddpm_weights,
                      default_latent, flatten_params, weights_sha256)
from .grid import axis_coords, grid_points, sign_change_mask
from .decoder import decoder_forward, decode_grid, decoder_forward_lowp, decoder_vjp_latent, decoder_vjp_latent_lowp, fit_loss_grad_lowp
from .ddpm import (ddpm_schedule, time_embedding, denoiser_forward,
                   ddpm_step, sample_latents, denoiser_forward_lowp, DDPM_LATENT_SCALE, to_decoder_latent)
from .philox import philox4x32_10, philox_normal_rows, philox_sampler_inputs
from .marching import marching_cubes, mesh_is_closed
from .train import (ddpm_train_grads, ddpm_train_grads_lowp, adam_step, flatten_grads, ddpm_noised_input,
                    decoder_train_grads, decoder_train_grads_lowp)
