#!/usr/bin/env python
"""bench.py - SDF decoder queries/s on N B200s (BASELINE.json metric), plus roofline,
end-to-end (host buffers through the C ABI) and a CPU-oracle baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path: decode_grid(z, 256) - BASELINE.json configs[1], one latent
on a 256^3 grid (16,777,216 queries), bf16 operands / fp32 accumulate, fused tcgen05 kernel.
With N > 1 every rank decodes its own latent's 256^3 grid (the path shards by independent
latents or z-slabs with no data-path collective; per-GPU work is fixed => weak scaling; the
per-rank query count equals one z-slab of configs[4]'s 512^3 grid on 8 GPUs).

`--impl reference`: the mounted reference has no source (/root/reference/README.md:1 is a
title), so the "reference arm" is the frozen CPU oracle (oracle/, a PyTorch fp32 restatement
of the method) timed on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = 256
QUERIES = RES ** 3
# work per query (SURVEY.md section 8d / DESIGN.md): MACs issued to the tensor pipe with the
# latent folded into biases and L3 padded to N=256: 6 * 512^2 ; dense count as the oracle computes it
FLOP_TENSOR_PER_QUERY = 2 * 6 * 512 * 512          # 3,145,728
FLOP_DENSE_PER_QUERY = 3_671_040
DDPM_LATENTS = 4096
DDPM_FLOP_PER_LATENT_STEP = 2 * (512 * 1024 + 3 * 1024 * 1024 + 1024 * 256)   # 7,864,320 executed (hi/lo split of x: K = 512)
NCU_DRAM_BYTES_PER_LAUNCH = 3400192 + 14253312      # profiles/r1_fused_decoder_ncu_full.csv (dram__bytes_read + write)
METRIC = "sdf_decoder_queries_per_s"
UNIT = "queries/s"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"burst": 1590.0, "sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                      # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:                   # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:                       # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_oracle_rate(target_s: float = 12.0):
    """Oracle (fp32 torch CPU) queries/s on a bounded sample: z-slabs of the 256^3 workload."""
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    z = oracle.default_latent()
    oracle.decode_grid(z, RES, 0, 1)                      # warm-up: one plane (65,536 queries)
    t0 = time.perf_counter()
    oracle.decode_grid(z, RES, 0, 2)
    dt = time.perf_counter() - t0
    planes = max(2, min(RES, int(target_s / max(dt / 2, 1e-6))))
    t0 = time.perf_counter()
    oracle.decode_grid(z, RES, 0, planes)
    dt = time.perf_counter() - t0
    q = planes * RES * RES
    return {"value": q / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"planes [0,{planes}) of the 256^3 grid = {q} queries in {dt:.2f} s, torch {torch.__version__} fp32, "
                      f"{cores} threads (oracle/decoder.py; no reference source exists to time)"}


def cpu_ddpm_rate(n: int = 2048, steps: int = 400):
    """Oracle DDPM sampler (fp32 torch CPU) latents/s on a bounded sample: n latents x `steps` of the 1000 steps."""
    import numpy as np
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rs = np.random.RandomState(0)
    x_T = rs.standard_normal((n, 256)).astype(np.float32)
    noise = rs.standard_normal((steps, n, 256)).astype(np.float32)
    oracle.sample_latents(n, x_T, noise[:5], steps=5)
    t0 = time.perf_counter()
    oracle.sample_latents(n, x_T, noise, steps=steps)
    dt = time.perf_counter() - t0
    return {"value": n / (dt * 1000.0 / steps), "unit": "latents/s", "cores": cores, "kind": "port",
            "sample": f"{n} latents x {steps} of 1000 steps in {dt:.2f} s (scaled to 1000 steps), torch fp32, {cores} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, args.steps), max(0, args.warmup)
    import torch
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    z = oracle.default_latent()
    planes = 4                                            # 262,144 queries per step (= configs[0]'s count)
    for _ in range(min(warm, 2)):
        oracle.decode_grid(z, RES, 0, planes)
    steps = min(steps, 12)
    t0 = time.perf_counter()
    for s in range(steps):
        oracle.decode_grid(z, RES, (s * planes) % RES, (s * planes) % RES + planes)
    dt = time.perf_counter() - t0
    q = planes * RES * RES
    value = q * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(warm, 2), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "decode_grid(z, 256): one latent, 256^3 grid (BASELINE configs[1]); each step a bounded "
                               f"sample of {planes} z-planes = {q} queries",
                   "note": "the mounted reference has no source; this arm is the frozen CPU oracle"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {q} queries, torch {torch.__version__} fp32, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the 512^3 z-slab-sharded case that runs when N > 1")
    ap.add_argument("--no-config4", action="store_true", help="skip the sample-then-decode leg (configs[3])")
    ap.add_argument("--no-ddpm", action="store_true", help="skip the latent-DDPM leg (second half of the metric)")
    ap.add_argument("--no-vjp", action="store_true", help="skip the latent-gradient leg (SURVEY 8f row N4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    K, W = max(1, args.steps), max(3, args.warmup)
    dev = torch.device("cuda", local)
    dec = pkg.Decoder(pkg.synthetic.decoder_params(), device=dev, precision=args.precision)
    z_host = pkg.synthetic.latent(rank)                            # each rank: its own latent (weak scaling)
    z = torch.from_numpy(z_host).to(dev)
    out = torch.empty((RES, RES, RES), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        dec.decode_grid(z, RES, out=out)
    barrier()
    kernel_ms, step_ms = [], []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(float(i))                                  # evict L2 between timed iterations (untimed)
            ev[i][0].record()
            dec.decode_grid(z, RES, out=out)                       # fold kernel + fused kernel on the current stream
            ev[i][1].record()
            ev[i][1].synchronize()
            kernel_ms.append(dec.last_kernel_ms())                 # events around the fused kernel itself
        barrier()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * QUERIES * K / (total_ms * 1e-3)

    # ---- end to end: numpy latent in, numpy sdf out through the host-buffer C-ABI call ----
    sdf_host = torch.empty((RES, RES, RES), dtype=torch.float32).pin_memory().numpy()
    for _ in range(2):
        dec.decode_grid_host(z_host, RES, out=sdf_host)
    barrier()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(Ke):
        dec.decode_grid_host(z_host, RES, out=sdf_host)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * QUERIES * Ke / float(e2e_s.item())
    checksum = float(np.float64(sdf_host[::16, ::16, ::16].sum()))

    # ---- BASELINE configs[4] (the north star's target case) when there is more than one GPU: ONE latent, 512^3 grid,
    # z-slab per rank, fused-path decode + sign-change mask (halo plane recomputed locally) + in-place NCCL all-gather
    cfg5 = None
    if world > 1 and not args.no_config5:
        try:
            res5 = 512
            z5 = torch.from_numpy(pkg.synthetic.latent(0)).to(dev)
            del out, flush
            for _ in range(2):
                s5, m5 = pkg.decode_grid_sharded(dec, z5, res5, mask=True)
            t5 = []
            for _ in range(5):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                s5, m5 = pkg.decode_grid_sharded(dec, z5, res5, mask=True)
                b.record()
                b.synchronize()
                t5.append(a.elapsed_time(b))
            t5 = torch.tensor([statistics.median(t5)], dtype=torch.float64, device=dev)
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            ms5 = float(t5.item())
            per_gpu_tflops = (res5 ** 3 / world) * FLOP_TENSOR_PER_QUERY / (ms5 * 1e-3) / 1e12
            cfg5 = {"workload": f"decode_grid_sharded(z, 512, mask=True) on {world} GPUs: z-slabs, mask with locally recomputed halo plane, "
                                "in-place NCCL all-gather of the sdf slabs and gather of the mask slabs; median of 5, max over ranks",
                    "ms": ms5, "queries_per_s": res5 ** 3 / (ms5 * 1e-3), "tflops_per_gpu_incl_mask_and_gather": per_gpu_tflops,
                    "active_cells": int(m5.sum().item()), "sdf_checksum": float(s5[::32, ::32, ::32].double().sum().item())}
            del s5, m5
        except Exception as exc:                     # an optional leg must never cost the headline line
            cfg5 = {"error": repr(exc)}
            print(f"bench.py: optional leg failed: {exc!r}", file=sys.stderr)

    # ---- second half of the metric: latent-DDPM latents/s (BASELINE configs[3]: 4096 latents, 1000 steps) ----
    ddpm_line = None
    if not args.no_ddpm:
        try:
            n_lat, T = DDPM_LATENTS, 1000
            sampler = pkg.LatentDDPM(pkg.synthetic.ddpm_params(), device=dev, precision=args.precision)
            g = torch.Generator(device=dev).manual_seed(100 + rank)
            x_T = torch.randn((n_lat, 256), generator=g, device=dev)
            noise = torch.randn((T, n_lat, 256), generator=g, device=dev)           # 4.2 GB explicit noise stream in HBM
            dd_ms = []
            for i in range(1 + 3):                                                   # 1 warm-up + 3 timed full samplings
                barrier()
                x0 = sampler.sample_latents(n_lat, x_T=x_T, noise=noise, steps=T)
                if i:
                    dd_ms.append(sampler.last_kernel_ms())
            dd = torch.tensor([statistics.mean(dd_ms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dd, op=dist.ReduceOp.MAX)
            dd_ms_max = float(dd.item())
            flop = n_lat * T * DDPM_FLOP_PER_LATENT_STEP
            ddpm_line = {"metric": "ddpm_latents_per_s", "value": world * n_lat / (dd_ms_max * 1e-3), "unit": "latents/s",
                         "workload": f"sample_latents({n_lat}) per GPU: MLP denoiser 4x1024, latent 256, {T} steps, explicit noise "
                                     "stream resident in HBM; ONE persistent cooperative kernel per call",
                         "ms_per_sampling": dd_ms_max, "us_per_step": dd_ms_max * 1e3 / T,
                         "kernel": "ddpm_sample_kernel", "achieved_tflops": flop / (dd_ms_max * 1e-3) / 1e12,
                         "flop_per_latent_step": DDPM_FLOP_PER_LATENT_STEP, "gpu_launches_per_sampling": 2,
                         "x0_abs_max": float(x0.abs().max().item())}
            # sample_latents(n) as the north star spells it - no stream argument: noise generated in the kernel
            # (Philox4x32-10), device-timed, then end to end through the host-buffer call (x_0 [n,256] comes back)
            sd_ms = []
            for i in range(1 + 3):
                barrier()
                xs0 = sampler.sample_latents(n_lat, steps=T, seed=7 + rank)
                if i:
                    sd_ms.append(sampler.last_kernel_ms())
            sd = torch.tensor([statistics.mean(sd_ms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(sd, op=dist.ReduceOp.MAX)
            ddpm_line["seeded"] = {"value": world * n_lat / (float(sd.item()) * 1e-3), "unit": "latents/s", "ms_per_sampling": float(sd.item()),
                                   "noise": "generated in the update epilogue (Philox4x32-10 + Box-Muller), no noise stream in memory",
                                   "x0_abs_max": float(xs0.abs().max().item())}
            sampler.sample_latents_seeded_host(n_lat, 7 + rank, steps=T)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                xh = sampler.sample_latents_seeded_host(n_lat, 7 + rank, steps=T)
            e2 = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(e2, op=dist.ReduceOp.MAX)
            ddpm_line["e2e"] = {"value": world * n_lat / float(e2.item()), "unit": "latents/s", "h2d_bytes_per_step": 0,
                                "d2h_bytes_per_step": int(xh.nbytes),
                                "api": "LatentDDPM.sample_latents_seeded_host -> sdfb_ddpm_sample_philox_host (x_T and noise generated on the device, x_0 returned to the host)"}
            del noise, sampler
        except Exception as exc:                     # an optional leg must never cost the headline line
            ddpm_line = {"error": repr(exc)}
            print(f"bench.py: optional leg failed: {exc!r}", file=sys.stderr)

    # ---- BASELINE configs[3], one GPU's share when it is spread over 8: 512 of the 4096 latents are sampled (1000 steps,
    # noise generated in the kernel) and each is decoded on a 128^3 grid (64 shapes per call into a reused buffer; 34 GB of
    # sdf in total for the full config, so the fields are consumed / discarded as they are produced)
    cfg4 = None
    if not args.no_config4:
        try:
            n4, res4 = 512, 128
            smp = pkg.LatentDDPM(pkg.synthetic.ddpm_params(), device=dev, precision=args.precision)
            buf4 = torch.empty((64, res4, res4, res4), dtype=torch.float32, device=dev)
            lat4 = smp.sample_latents(n4, seed=1000 + rank)
            dec.decode_grid_batch(lat4[:64], res4, out=buf4)
            barrier()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            lat4 = smp.sample_latents(n4, seed=1000 + rank)
            b.record()
            for i0 in range(0, n4, 64):
                dec.decode_grid_batch(lat4[i0:i0 + 64], res4, out=buf4)
            c.record()
            c.synchronize()
            t4 = torch.tensor([a.elapsed_time(b), b.elapsed_time(c)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t4, op=dist.ReduceOp.MAX)
            q4 = n4 * res4 ** 3
            cfg4 = {"workload": f"per GPU: sample_latents({n4}) (1000 steps, in-kernel noise) then decode_grid_batch of the {n4} samples at "
                                f"{res4}^3 (64 per call); {world} GPUs cover {world * n4} latents (BASELINE configs[3] = 4096 latents on 8 GPUs)",
                    "ddpm_ms": float(t4[0].item()), "decode_ms": float(t4[1].item()), "total_ms": float(t4.sum().item()),
                    "decode_queries_per_s": world * q4 / (float(t4[1].item()) * 1e-3),
                    "decode_tflops_per_gpu": q4 * FLOP_TENSOR_PER_QUERY / (float(t4[1].item()) * 1e-3) / 1e12}
            del buf4, smp, lat4
        except Exception as exc:                     # an optional leg must never cost the headline line
            cfg4 = {"error": repr(exc)}
            print(f"bench.py: optional leg failed: {exc!r}", file=sys.stderr)

    # ---- latent gradient (auto-decoder fitting) through the forward + backward instance of the fused
    # kernel, next to the fp32 path on a sample of the same points
    vjp = None
    if not args.no_vjp:
        try:
            n_v = 1 << 22
            gv = torch.Generator(device=dev).manual_seed(7 + rank)
            pts = torch.rand((n_v, 3), generator=gv, device=dev) * 2 - 1
            up = torch.randn(n_v, generator=gv, device=dev) / n_v
            zv = torch.from_numpy(pkg.synthetic.latent(rank)).to(dev)
            dec.latent_vjp(zv, pts, up, precision=args.precision)
            barrier()
            ms_v = []
            for _ in range(3):
                dec.latent_vjp(zv, pts, up, precision=args.precision)
                ms_v.append(dec.last_kernel_ms())
            n32 = 1 << 17
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dec.latent_vjp(zv, pts[:n32], up[:n32], precision="fp32")
            a.record()
            dec.latent_vjp(zv, pts[:n32], up[:n32], precision="fp32")
            b.record()
            b.synchronize()
            t_v = torch.tensor([statistics.median(ms_v)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_v, op=dist.ReduceOp.MAX)
            kv = float(t_v.item())
            vjp = {"workload": f"per GPU: latent_vjp over {n_v} random points, precision {args.precision} (one launch of "
                               "fused_decoder_kernel<., BWD>: forward + 13 backward passes per tile)",
                   "kernel_ms": kv, "points_per_s": world * n_v / (kv * 1e-3),
                   "achieved_tflops_per_gpu": n_v * 2 * FLOP_TENSOR_PER_QUERY / (kv * 1e-3) / 1e12,
                   "fp32_path_points_per_s": n32 / (a.elapsed_time(b) * 1e-3)}
            del pts, up
        except Exception as exc:                     # an optional leg must never cost the headline line
            vjp = {"error": repr(exc)}
            print(f"bench.py: optional leg failed: {exc!r}", file=sys.stderr)

    if rank == 0:
        peaks = read_peaks()
        k_ms = statistics.mean(kernel_ms)
        achieved = QUERIES * FLOP_TENSOR_PER_QUERY / (k_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "decode_grid(z, 256): one latent per GPU, 256^3 grid = 16,777,216 queries "
                                   "(BASELINE configs[1]); seeded random-init decoder 8x512, latent 256",
                       "l2": "256 MiB buffer written between timed iterations (L2 flush, untimed)",
                       "timing": "CUDA events per step on the launching stream, summed, max over ranks",
                       "parallelism": f"dp{world} (independent latents, no data-path collective)"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 1024, "d2h_bytes_per_step": QUERIES * 4,
                    "steps": Ke, "api": "Decoder.decode_grid_host -> sdfb_decode_grid_host (numpy in, pinned numpy out)",
                    "checksum": checksum},
            "gpu_launches": 2 * K,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["burst"], "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture "
                                           "profiles/r1_fused_decoder_ncu_full.csv (not re-measured in this run); algorithmic "
                                           "bytes = 67,108,864 output bytes, part of which is still in L2 when the launch ends",
                         "peak_kind": "burst bf16 matmul, " + peaks["source"],
                         "frac_of_sustained": achieved / peaks["sustained"], "peak_sustained": peaks["sustained"],
                         "kernel": "fused_decoder_kernel", "kernel_ms": k_ms,
                         "flop_per_query_tensor_pipe": FLOP_TENSOR_PER_QUERY,
                         "dense_equiv_tflops": QUERIES * FLOP_DENSE_PER_QUERY / (k_ms * 1e-3) / 1e12},
            "wall_s_timed_region": t_wall,
        }
        if cfg5 is not None:
            if "error" not in cfg5:
                cfg5["frac_of_burst_peak_per_gpu"] = cfg5["tflops_per_gpu_incl_mask_and_gather"] / peaks["burst"]
            line["config5_512cubed_sharded"] = cfg5
        if cfg4 is not None:
            line["config4_sample_then_decode"] = cfg4
        if vjp is not None:
            if "achieved_tflops_per_gpu" in vjp:
                vjp["frac_of_burst_peak"] = vjp["achieved_tflops_per_gpu"] / peaks["burst"]
            line["latent_gradient"] = vjp
        if ddpm_line is not None:
            if "achieved_tflops" in ddpm_line:
                ddpm_line["frac_of_burst_peak"] = ddpm_line["achieved_tflops"] / peaks["burst"]
            line["ddpm"] = ddpm_line
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_rate()
            if ddpm_line is not None and "error" not in ddpm_line:
                line["ddpm"]["cpu_baseline"] = cpu_ddpm_rate()
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
