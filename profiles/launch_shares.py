#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
Usage: python profiles/launch_shares.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<name>.txt"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iN, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows:
    if r is hdr or r[iM] != "gpu__time_duration.sum":
        continue
    us = float(r[iV].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1e-3)
    name = re.sub(r"\(.*$", "", r[iN]).replace("void ", "").replace("mgr::", "").strip()
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += us
all_us = sum(v[1] for v in tot.values())
print(f"# Launch list summary (ncu --metrics gpu__time_duration.sum --clock-control none, {sys.argv[2] if len(sys.argv) > 2 else ''})")
print("# cold-cache, serialised: compare SHARES, not absolutes.\n")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    if us / all_us > 0.002:
        print(f"{n:4d} launches  {us:10.1f} us total  {us / n:9.1f} us/launch  {100 * us / all_us:5.1f}%  {name}")
