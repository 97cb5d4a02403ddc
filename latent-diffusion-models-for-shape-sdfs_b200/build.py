"""Builds libsdfb200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object, see include/sdfb200.h)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("SDFB_LIB") or os.path.join(HERE, "libsdfb200.so")     # SDFB_LIB: an experimental build beside the product one
SOURCES = ("api.cu", "fused_decoder.cu", "tc_common.cu", "fp32_kernels.cu", "umma_rate.cu", "tma_ingest.cu", "ddpm_step.cu", "marching.cu", "comm.cu", "sparse.cu", "gemm_tc.cu", "train_kernels.cu", "train_api.cu")
HEADERS = ("kernels.h", "ptx.cuh", "philox.cuh", "mc_tables.h", "comm.h", os.path.join("..", "..", "include", "sdfb200.h"))
NVCC_FLAGS = ("-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsdfb200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libsdfb200.so (skipped when up to date). Returns the path."""
    if not force and not is_stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("SDFB_NVCC_EXTRA", "").split()      # e.g. -DSDFB_K1_EXPERIMENTS for the timing experiments
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB + ".tmp", *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
