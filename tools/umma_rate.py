#!/usr/bin/env python
"""Dev tool: cycles per tcgen05.mma for a few issue patterns (see csrc/umma_rate.cu)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
from __graft_entry__ import load_package  # noqa: E402
lib = load_package().load_library()
torch.zeros(1, device="cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else 'ss'
if mode == 'ss':
    for cg in (1, 2):
        for kpc in (1, 2, 4, 8):
            for flags in (0, 1, 2, 3, 5, 7):
                v = C.c_double()
                rc = lib.sdfb_umma_rate(cg, 148, 200, kpc, 2, flags, C.byref(v))
                print(f"cg={cg} k_per_commit={kpc:2d} flags={flags} (nowait={flags&1} always_acc={(flags>>1)&1} 2commits={(flags>>2)&1}): rc={rc} {v.value:7.1f} cycles/MMA", flush=True)
if mode == 'epi':
    for flags, name in ((16, "tcgen05.st + wait::st + fence"), (32, "st.shared x4 + fence.proxy.async")):
        v = C.c_double()
        rc = lib.sdfb_umma_rate(1, 148, 2000, 1, 2, flags, C.byref(v))
        print(f"epilogue-shaped loop, hand-over by {name}: rc={rc} {v.value * 4:7.1f} cycles per 32-column chunk per warp", flush=True)
if mode == 'ts':
    # TMEM-operand probes (DESIGN §3 "Round 2"): the TS form an activations-in-TMEM decoder would issue (N = 128), and the
    # epilogue hand-over cost through tcgen05.st against the shipped st.shared + proxy fence
    extra = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 64: N = 256; 128: operand below the accumulators
    grid = int(sys.argv[3]) if len(sys.argv) > 3 else 148
    n_acc = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    base = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    for cg in (1, 2):
        for kpc in (4, 8, 16):
            for flags in ((base + 1,) if not (extra & 64) else (base, base + 1)):   # N = 128 only without intermediate waits: the
                # one-outstanding-group protocol of the probe assumes the pipe is slower than the issuing thread
                v = C.c_double()
                print(f"TS form (A in TMEM) base={base} n_acc={n_acc} extra={extra} grid={grid} cg={cg} k_per_commit={kpc:2d} nowait={flags&1}: ", end="", flush=True)
                rc = lib.sdfb_umma_rate(cg, grid, 200, kpc, n_acc, flags | extra, C.byref(v))
                print(f"rc={rc} {v.value:7.1f} cycles/MMA", flush=True)
