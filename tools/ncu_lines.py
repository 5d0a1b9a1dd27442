#!/usr/bin/env python
"""Per-CUDA-source-line view of an Nsight Compute capture (developer tool): where the warp-instructions and the stall
samples of one kernel go, by source line and by inlined function body.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> > src.csv
    python tools/ncu_lines.py src.csv [top=40] [launch=0]

(needs -lineinfo at compile time and --import-source on at capture time; the csv holds one block per source file per
profiled launch -- `launch` picks the n-th launch.)"""
import csv
import os
import sys
from collections import defaultdict

path = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0

rows = csv.reader(open(path, newline=""))
blocks, cur = [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = {"file": r[1], "func": None, "hdr": None, "lines": []}
        blocks.append(cur)
    elif r[0] == "Function Name" and cur is not None:
        cur["func"] = r[1]
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r[0].isdigit():
        cur["lines"].append(r)

# the n-th launch = the n-th occurrence of each file within the capture
seen, pick = defaultdict(int), []
for b in blocks:
    k = (b["file"], b["func"])
    if seen[k] == launch:
        pick.append(b)
    seen[k] += 1
if not pick:
    sys.exit("no blocks")
print("kernel:", (pick[0]["func"] or "")[:140])
tot_i = tot_s = 0
items = []
per_file = defaultdict(lambda: [0, 0])
for b in pick:
    h = b["hdr"]
    iS, iI = h.index("# Samples"), h.index("Instructions Executed")
    for r in b["lines"]:
        try:
            s, i = int(r[iS]), int(r[iI])
        except ValueError:
            continue
        tot_i += i
        tot_s += s
        per_file[os.path.basename(b["file"])][0] += i
        per_file[os.path.basename(b["file"])][1] += s
        items.append((i, s, os.path.basename(b["file"]), int(r[0]), r[1].strip()[:110]))
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for f, (i, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} instr {100 * i / tot_i:5.1f}%  samples {100 * s / max(tot_s, 1):5.1f}%")
print("-- top lines by executed warp-instructions")
for i, s, f, ln, src in sorted(items, key=lambda t: -t[0])[:ntop]:
    print(f"  {100 * i / tot_i:5.1f}% i {100 * s / max(tot_s, 1):5.1f}% s  {f}:{ln:<4d} {src}")
print("-- top lines by stall samples")
for i, s, f, ln, src in sorted(items, key=lambda t: -t[1])[:ntop]:
    print(f"  {100 * s / max(tot_s, 1):5.1f}% s {100 * i / tot_i:5.1f}% i  {f}:{ln:<4d} {src}")
