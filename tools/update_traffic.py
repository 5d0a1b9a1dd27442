#!/usr/bin/env python
"""Write profiles/roofline_traffic.json from an `ncu --set full` report: DRAM bytes (read + write) per
launch of the dominant kernel of the backward (pass 1 + pass 2 summed, one launch each per step).
Usage: python tools/update_traffic.py gpurun_out/prof.ncu-rep c2"""
import csv
import json
import os
import subprocess
import sys

rep, workload = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


per_kernel = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("mgr::", "")
    b = sum(to_bytes(d[k], units[hdr.index(k)]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per_kernel.setdefault(name, []).append(b)
avg = {k: sum(v) / len(v) for k, v in per_kernel.items()}
bwd = sum(v for k, v in avg.items() if "bwd" in k)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[workload] = {"bytes": bwd, "per_kernel": avg, "source": f"ncu --set full, {os.path.basename(rep)}: dram__bytes_read.sum + dram__bytes_write.sum per launch"}
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[workload], indent=1))
