#!/usr/bin/env python
"""Write profiles/roofline_traffic.json from an `ncu --set full` report of `bench.py --kernels-only`: DRAM bytes (read +
write) per launch of every kernel, averaged over its ACTIVE launches (a kernel whose samples belong to the other path
leaves after one look at the placements: a few KB), and summed into the figures bench.py reports:

    bytes                 backward, general placements   (placement kernels + pass 1 + pass 2)
    forward               forward, general placements
    translation_fwd_bwd   forward + backward of the translation-only leg (stencil kernels)

Usage: python tools/update_traffic.py gpurun_out/prof_general.ncu-rep [gpurun_out/prof_translation.ncu-rep ...] c2"""
import csv
import json
import os
import subprocess
import sys

reps, workload = sys.argv[1:-1], sys.argv[-1]
rep = "+".join(os.path.basename(r) for r in reps)
rows = []
for r_ in reps:
    txt = subprocess.run(["ncu", "-i", r_, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    hdr, units = rr[0], rr[1]
    rows += [dict(zip(hdr, r)) for r in rr[2:]]


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


per_kernel = {}
for d in rows:
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("mgr::", "")
    b = sum(to_bytes(d[k], units[hdr.index(k)]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    per_kernel.setdefault(name, []).append(b)
# active launches of a kernel: within a factor 4 of its largest launch
avg = {}
for k, v in per_kernel.items():
    act = [b for b in v if b * 4 >= max(v)]
    avg[k] = sum(act) / len(act)


def total(*keys):
    return sum(v for k, v in avg.items() if any(s in k for s in keys))


path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[workload] = {
    "bytes": total("render_bwd_pass1", "render_bwd_pass2", "placements", "inverse_plans", "sample_flags"),
    "forward": total("render_fwd_ws", "render_fwd_general_only", "render_fwd<"),
    "translation_fwd_bwd": total("render_fwd_stencil_only", "render_fwd_shift_tma", "render_bwd_shift"),
    "per_kernel": avg,
    "source": f"ncu --set full, {rep}: dram__bytes_read.sum + dram__bytes_write.sum per active launch"}
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[workload], indent=1))
