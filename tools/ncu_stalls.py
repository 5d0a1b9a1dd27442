#!/usr/bin/env python
"""Stall-reason totals and the hottest SASS lines of one kernel from an Nsight Compute capture (developer tool).
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_stalls.py src.csv [top=25]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], newline="")))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
best = None
for k, i in enumerate(hdr_idx):
    j = hdr_idx[k + 1] if k + 1 < len(hdr_idx) else len(rows)
    if best is None or j - i > best[1] - best[0]:
        best = (i, j)
i, j = best
h = rows[i]
cols = {n: c for c, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
S = {n: 0 for n in stall_cols}
samples = instr = 0
data = []
for r in rows[i + 1:j]:
    if len(r) < len(h):
        continue
    try:
        ns = int(r[cols["# Samples"]])
    except ValueError:
        continue
    samples += ns
    instr += int(r[cols["Instructions Executed"]] or 0)
    for n in stall_cols:
        S[n] += int(r[cols[n]] or 0)
    data.append((ns, r))
print("samples", samples, "warp-instructions", instr)
for n, v in sorted(S.items(), key=lambda x: -x[1])[:10]:
    print(f"  {n:24s} {v:7d} {100 * v / max(samples, 1):5.1f}%")
data.sort(key=lambda x: -x[0])
for ns, r in data[:ntop]:
    st = {n: int(r[cols[n]] or 0) for n in stall_cols}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{ns:6d} {r[3][:72]:72s} {top}")
