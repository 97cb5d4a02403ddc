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
for cg in (1, 2):
    for kpc in (1, 2, 4, 8):
        for flags in (0, 1, 2, 3, 5, 7):
            v = C.c_double()
            rc = lib.sdfb_umma_rate(cg, 148, 200, kpc, 2, flags, C.byref(v))
            print(f"cg={cg} k_per_commit={kpc:2d} flags={flags} (nowait={flags&1} always_acc={(flags>>1)&1} 2commits={(flags>>2)&1}): rc={rc} {v.value:7.1f} cycles/MMA")
