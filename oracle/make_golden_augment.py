"""Generate tests/golden/augment_geom_golden.npz by running the REAL reference's AugmentPipe (geometric transforms only,
read-only import from /root/reference/montage_gan) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container:

    python oracle/make_golden_augment.py

The pipe draws its transform inside ``forward``; with ``debug_percentile`` the draw is deterministic.  What the block
does with it is observed from outside: the reflect padding it asks of ``F.pad``, the 2x3 matrices it hands to
``F.affine_grid`` and its output.  ``G_inv`` itself is recovered from those (``oracle.augment_geom.recover_G_inv``);
the recovered matrix must then reproduce the OBSERVED padding through the restatement's own corner arithmetic, which is
what makes the recovery a check and not a tautology.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import augment_geom as AG  # noqa: E402

REF = "/root/reference/montage_gan"


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, REF)
    from training.augment import AugmentPipe  # noqa: E402  (the reference, unmodified)
    seen = {}
    orig_ag, orig_pad = F.affine_grid, F.pad

    def spy_affine_grid(theta, size, align_corners=None):
        seen["theta"], seen["size"] = theta.detach().clone(), [int(v) for v in size]
        return orig_ag(theta, size, align_corners=align_corners)

    def spy_pad(input, pad, mode="constant", value=None):
        if mode == "reflect":
            seen["pad"] = [int(p) for p in pad]
        return orig_pad(input, pad, mode=mode) if value is None else orig_pad(input, pad, mode=mode, value=value)

    F.affine_grid = torch.nn.functional.affine_grid = spy_affine_grid
    F.pad = torch.nn.functional.pad = spy_pad
    out, names = {}, []
    cases = [("all_p30", dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1), 0.30, (2, 4, 32, 32)),
             ("all_p80", dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1), 0.80, (2, 4, 48, 40)),
             ("rotate_scale_p65", dict(scale=1, rotate=1), 0.65, (1, 4, 64, 64)),
             ("xfrac_only_p10", dict(xfrac=1), 0.10, (2, 3, 24, 24)),
             ("aniso_p95", dict(aniso=1, rotate=1), 0.95, (1, 4, 40, 56))]
    g = torch.Generator().manual_seed(0)
    for name, kw, pct, shape in cases:
        pipe = AugmentPipe(**kw)
        pipe.p.copy_(torch.as_tensor(1.0))
        x = torch.rand(shape, generator=g) * 2 - 1
        seen.clear()
        y = pipe(x, debug_percentile=torch.as_tensor(pct))
        mx0, mx1, my0, my1 = seen["pad"]
        H, W = shape[2], shape[3]
        G_inv = AG.recover_G_inv(seen["theta"], H, W, mx0, my0, mx1, my1)
        out[f"{name}/images"] = x.numpy()
        out[f"{name}/G_inv"] = G_inv.numpy()
        out[f"{name}/margins"] = np.array([mx0, my0, mx1, my1])
        out[f"{name}/theta"] = seen["theta"].numpy()
        out[f"{name}/grid_size"] = np.array(seen["size"])
        out[f"{name}/out"] = y.detach().numpy()
        names.append(name)
        print(name, shape, "pad", seen["pad"], "grid", seen["size"], "mean |out|", float(y.abs().mean()))
    F.affine_grid = torch.nn.functional.affine_grid = orig_ag
    F.pad = torch.nn.functional.pad = orig_pad
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "augment_geom_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
