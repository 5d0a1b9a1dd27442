#!/usr/bin/env python
"""Where the warp-time goes: stall samples and executed instructions of an `ncu --page source --csv` dump in chunks
of N SASS lines, plus the individual instructions with the most samples (developer tool).
usage: sass_samples.py <src.csv> [chunk=100] [top=30]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iSm]) for r in data)
ti = sum(int(r[iI]) for r in data)
print("total samples", tot, "warp instr", ti, "SASS lines", len(data))
for a in range(0, len(data), chunk):
    sm = sum(int(r[iSm]) for r in data[a:a + chunk])
    ins = sum(int(r[iI]) for r in data[a:a + chunk])
    print(f"{a:5d} samples {100 * sm / tot:5.1f}%  instr {100 * ins / ti:5.1f}%  exec[first]={data[a][iI]}")
top = sorted(range(len(data)), key=lambda k: -int(data[k][iSm]))[:ntop]
for k in sorted(top):
    print(k, data[k][iSm], data[k][iI], data[k][iS][:90])
