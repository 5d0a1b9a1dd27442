#!/usr/bin/env python
"""A/B kernel variants without paying a rebuild on the GPU box (developer tool).

    here:        python tools/ab.py build <name> "-DMGR_FOO=3 -DMGR_BAR=1"     # -> tools/variants/<name>.so (+ .hash, .defs)
    on the box:  python tools/ab.py run <name> -- python tools/kbench.py ...   # swaps the variant in, runs, restores

Variants are selected with -D defines only (MGR_NVCC_DEFINES, see build.py), so the sources -- and therefore the
staleness hash -- are the same for all of them apart from the recorded defines."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "docker-montage-gan_b200")
LIB = os.path.join(PKG, "libmontage_render.so")
VAR = os.path.join(ROOT, "tools", "variants")


def main():
    mode, name = sys.argv[1], sys.argv[2]
    os.makedirs(VAR, exist_ok=True)
    base = os.path.join(VAR, name)
    if mode == "build":
        defs = sys.argv[3] if len(sys.argv) > 3 else ""
        keep = {p: open(p, "rb").read() for p in (LIB, LIB + ".hash") if os.path.isfile(p)}
        env = dict(os.environ, MGR_NVCC_DEFINES=defs)
        subprocess.run([sys.executable, os.path.join(PKG, "build.py"), "--force"], env=env, check=True, stdout=subprocess.DEVNULL)
        shutil.copy(LIB, base + ".so")
        shutil.copy(LIB + ".hash", base + ".hash")
        open(base + ".defs", "w").write(defs)
        for p, data in keep.items():
            open(p, "wb").write(data)
        print("built", base + ".so", "with", defs or "(no defines)")
    elif mode == "run":
        cmd = sys.argv[sys.argv.index("--") + 1:]
        keep = {p: open(p, "rb").read() for p in (LIB, LIB + ".hash") if os.path.isfile(p)}
        shutil.copy(base + ".so", LIB)
        shutil.copy(base + ".hash", LIB + ".hash")
        env = dict(os.environ, MGR_NVCC_DEFINES=open(base + ".defs").read())
        try:
            print(f"== variant {name}: {env['MGR_NVCC_DEFINES']}", flush=True)
            rc = subprocess.run(cmd, env=env).returncode
        finally:
            for p, data in keep.items():
                open(p, "wb").write(data)
        sys.exit(rc)


if __name__ == "__main__":
    main()
