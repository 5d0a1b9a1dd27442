#!/usr/bin/env python
"""Static view of a kernel's SASS: total instruction count and every backward branch (loop) with its body
length and opcode mix.  usage: sass_loops.py <lib.so> <substring of the mangled kernel name> [...more substrings]
(developer tool; complements sass_regions.py, which needs an ncu capture)."""
import re
import subprocess
import sys

lib, keys = sys.argv[1], sys.argv[2:]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if not all(k in name for k in keys):
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, "instructions:", len(ins))
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    for k, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*\s.*?0x([0-9a-f]+)", t)
        m2 = re.search(r"0x([0-9a-f]+)", t) if "BRA" in t else None
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = ins[addr_index[tgt]:k + 1]
                ops = {}
                for _, tt in body:
                    tok = tt.split()
                    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
                top = sorted(ops.items(), key=lambda kv: -kv[1])[:10]
                print(f"  loop {addr_index[tgt]:5d}..{k:5d}  len {len(body):4d}  {top}")
