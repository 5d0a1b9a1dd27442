"""Ahead-of-time build of libmontage_render.so (sm_100a) with nvcc -- no torch headers, no JIT.

The reference JIT-builds its plugins through ``torch.utils.cpp_extension.load`` with an md5 cache
(``torch_utils/custom_ops.py:49-129``); here the library is built in-tree once (nvcc cross-compiles
without a GPU) and travels to the GPU box.  A content hash of the sources decides staleness, not
mtimes, so a copied tree does not rebuild.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libmontage_render.so")
HASH_PATH = LIB_PATH + ".hash"
SOURCES = ["mgr_api.cu", "inst_f32.cu", "inst_bf16.cu", "inst_f16.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
] + (["-DMGR_EXPERIMENT_NO_STAGE_LOADS"] if os.environ.get("MGR_EXPERIMENT_NO_STAGE_LOADS") else []) \
  + os.environ.get("MGR_NVCC_DEFINES", "").split()      # developer knob for A/B builds (tools/ab.py); part of the hash


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for f in files:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(HASH_PATH)):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libmontage_render.so if the sources changed.  Returns the path."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", "-o", obj, os.path.join(CSRC, src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + proc.stdout + proc.stderr, proc.returncode

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = "\n".join(r[1] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB_PATH] + [r[0] for r in results]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += "\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        rc = proc.returncode
    with open(os.path.join(PKG_DIR, "build.log"), "w") as fh:
        fh.write(log)
    if rc != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    with open(HASH_PATH, "w") as fh:
        fh.write(source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
