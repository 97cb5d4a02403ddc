#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_lines.py src.csv [top_n] [kernel_index]
Only the first kernel (launch) in the report is counted unless kernel_index is given."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
per_line = defaultdict(lambda: [0, defaultdict(int)])
hdr = None
file_name = "?"
seen_kernels = []
cur_line = None
total = 0
kernel_idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kcount = -1
active = False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        file_name = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Kernel Name":
        kcount += 1
        active = kcount == kernel_idx
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        if kcount < 0:
            active = True
        continue
    if hdr is None or not active:
        continue
    if r[0] != "":
        cur_line = (file_name, r[0], r[1].strip()[:90])
        continue
    # sass row belonging to cur_line
    try:
        s = int(r[i_samp])
    except (ValueError, IndexError):
        continue
    total += s
    per_line[cur_line][0] += s
    for i, name in stall_cols:
        if i < len(r) and r[i] not in ("", "0"):
            per_line[cur_line][1][name] += int(r[i])
print(f"total samples {total}")
for key, (s, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k}:{100 * v / max(s, 1):.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{100 * s / max(total, 1):5.1f}% {key[0]}:{key[1]:>4} {key[2]:90s} [{tops}]")
