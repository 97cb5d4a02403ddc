#!/usr/bin/env python
"""Dev tool: L2 -> shared memory operand delivery rate through the TMA unit (see csrc/tma_ingest.cu)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
from __graft_entry__ import load_package  # noqa: E402
lib = load_package().load_library()
torch.zeros(1, device="cuda")
NAMES = {0: "own boxes", 1: "cluster loads the same boxes", 2: "multicast to the cluster", 3: "1-D bulk 16 KiB", 4: "1-D bulk 4 x 4 KiB",
         5: "pair loads on the leader's barrier", 6: "pair, own barriers + forward"}


def run(grid, cluster, mode, issuers, uniform, cols=64, mib=32, iters=4096):
    out = (C.c_double * 3)()
    rc = lib.sdfb_tma_ingest_rate(grid, cluster, mode, issuers, uniform, cols, mib, iters, out)
    how = "elected lane of a uniform warp" if uniform else "lane 0 alone in its loop"
    print(f"grid={grid:3d} cluster={cluster} mode={mode} ({NAMES[mode]:28s}) issuing warps={issuers} ({how:30s}) cols={cols:4d} tensor={mib:4d} MiB: "
          f"rc={rc} {out[0]:6.1f} B/clk/CTA mean, {out[1]:6.1f} slowest, {out[2]:8.1f} GB/s aggregate", flush=True)


for uniform in (0, 1):
    for issuers in (1, 2, 4):
        for mode in (0, 3):
            for grid in (1, 148):
                run(grid, 1, mode, issuers, uniform)
run(148, 1, 0, 1, 1, cols=2560)
run(148, 1, 0, 2, 1, cols=2560)
for cluster in (2, 8):
    for mode in (1, 2):
        run(128, cluster, mode, 1, 1)
        run(128, cluster, mode, 2, 1)
for issuers in (1, 2, 4):
    run(128, 2, 5, issuers, 0)
    run(128, 2, 6, issuers, 0)
    run(128, 2, 0, issuers, 0)
run(148, 1, 0, 2, 1, mib=4096, iters=1024)       # not L2-resident: every CTA streams its own 16 MiB of a 4 GiB tensor from HBM
run(148, 1, 3, 2, 1, mib=4096, iters=1024)
