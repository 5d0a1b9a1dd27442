#!/usr/bin/env python
"""Group the SASS of an `ncu --page source --csv` dump into regions of equal execution count and
print each region's share of executed warp-instructions and its opcode mix (developer tool)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hdr, data = rows[1], rows[2:]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iI]) for r in data)
print("total warp instr", tot, "SASS lines", len(data))
regions, prev = [], None
for k, r in enumerate(data):
    c = int(r[iI])
    if prev is None or abs(c - prev) > 0.02 * max(c, prev, 1):
        regions.append([k, k, c, 0, 0])
    regions[-1][1] = k
    regions[-1][3] += c
    regions[-1][4] += int(r[iSm])
    prev = c
for a, b, c, s, sm in regions:
    if s > thresh * tot:
        ops = {}
        for r in data[a:b + 1]:
            tok = r[iS].split()
            op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:14]
        print(f"instr {a:4d}-{b:4d} n={b - a + 1:4d} exec/instr={c:9d} share={100 * s / tot:5.1f}% samples={sm:5d}  {top}")
