"""Generate tests/golden/pil_composite_golden.npz by running the REAL reference's Pillow composite
(``custom_utils.image_utils.alpha_composite``, read-only import from /root/reference/montage_gan) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not travel to the GPU box; the
vectors do):

    python oracle/make_golden_pil.py

Pins ``oracle/restatement.py: pil_alpha_composite`` and the CUDA kernel behind ``mgr_composite_u8`` bit for bit.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import montage_gan_b200  # noqa: E402,F401
from montage_gan_b200 import synth  # noqa: E402
from oracle import torch_chain as TC  # noqa: E402


def edge_alpha_case(seed):
    """every (dst alpha, src alpha) byte pair once, random colours: [1, 2, 4, 256, 256]"""
    g = np.random.default_rng(seed)
    x = g.integers(0, 256, (1, 2, 4, 256, 256)).astype(np.float32)
    x[0, 0, 3] = np.arange(256, dtype=np.float32)[:, None]
    x[0, 1, 3] = np.arange(256, dtype=np.float32)[None, :]
    # byte b is reproduced by trunc(v * 255) for v = (b + 0.5) / 255
    return torch.from_numpy(((x + 0.5) / 255.0).astype(np.float32))


def main():
    iu, _ = TC.load_reference()
    out, names = {}, []
    cases = [("smooth_L7", synth.make_layers(2, 7, 32, 32, "S", seed=300), "m11"),
             ("noise_L4", synth.make_layers(2, 4, 24, 20, "W", seed=301), "m11"),
             ("sparse_L9", synth.make_layers(1, 9, 32, 32, "F", seed=302), "m11"),
             ("single_layer", synth.make_layers(2, 1, 8, 8, "F", seed=303), "m11"),
             ("range01", (synth.make_layers(1, 5, 16, 16, "W", seed=304) + 1) / 2, "01"),
             ("all_alpha_pairs", edge_alpha_case(305), "01")]
    for name, x, in_range in cases:
        z = iu.normalize_zero1(x) if in_range == "m11" else x
        ref = iu.alpha_composite(z)                           # [B,4,H,W] fp32 in [0,1]
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/in_range"] = np.array(in_range)
        out[f"{name}/out"] = ref.numpy()
        names.append(name)
        print(name, tuple(x.shape), in_range, "mean", float(ref.mean()))
    unb = synth.make_layers(1, 3, 8, 8, "W", seed=306)[0]
    out["unbatched/x"] = unb.numpy()
    out["unbatched/in_range"] = np.array("m11")
    out["unbatched/out"] = iu.alpha_composite(iu.normalize_zero1(unb)).numpy()
    names.append("unbatched")
    out["names"] = np.array(names)
    import PIL
    import torchvision
    out["versions"] = np.array(f"Pillow {PIL.__version__}, torchvision {torchvision.__version__}, torch {torch.__version__}")
    path = os.path.join(ROOT, "tests", "golden", "pil_composite_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
