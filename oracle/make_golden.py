"""Generate tests/golden/render_golden.npz by running the REAL reference (read-only import from
/root/reference/montage_gan) on CPU in fp32 and fp64.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not travel to the
GPU box; the vectors do):

    python oracle/make_golden.py

The reference publishes no golden vectors for this path, so these outputs of
the reference's own code are what pins both ``oracle/restatement.py`` and the CUDA kernels.
Functions exercised: ``custom_utils.image_utils.{alpha_composite_pytorch, normalize_zero1,
normalize_minus11, convert_translate_to_2x3}`` and the STNv2c warp lines
(``fukuwarai/networks.py:250-257``) via ``torch.nn.functional.{affine_grid, grid_sample}``.
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

CASES = [
    # name, B, L, H, W, layer family, theta family (None = composite only), cover, in_range, grad_out
    ("affine_small", 2, 3, 12, 10, "W", "I", True, "m11", "randn"),
    ("affine_smooth", 1, 4, 32, 32, "S", "I", True, "m11", "ones"),
    ("translate", 1, 7, 16, 16, "S", "T", True, "m11", "randn"),
    ("extreme", 2, 4, 20, 24, "W", "X", True, "m11", "randn"),
    ("sparse", 2, 5, 16, 16, "F", "T", False, "m11", "randn"),
    ("identity", 1, 3, 8, 8, "W", "0", False, "m11", "randn"),
    ("composite_only", 2, 4, 8, 8, "W", None, False, "m11", "randn"),
    ("composite_only_sparse", 1, 9, 16, 16, "F", None, False, "m11", "randn"),
    ("range01", 1, 3, 8, 8, "W", "I", True, "01", "randn"),
    ("single_layer", 2, 1, 8, 8, "F", "T", False, "m11", "randn"),
]


def main():
    iu, _ = TC.load_reference()
    out = {}
    names = []
    for i, (name, B, L, H, W, lf, tf, cover, in_range, gk) in enumerate(CASES):
        x = synth.make_layers(B, L, H, W, lf, seed=100 + i)
        if in_range == "01":
            x = (x + 1) / 2
        theta = None if tf is None else synth.make_theta(B, L, tf, seed=100 + i, cover_back=cover)
        go = synth.make_grad_out(B, H, W, gk, seed=100 + i)
        out[f"{name}/x"] = x.numpy()
        if theta is not None:
            out[f"{name}/theta"] = theta.numpy()
        out[f"{name}/grad_out"] = go.numpy()
        out[f"{name}/in_range"] = np.array(in_range)
        for tag, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
            r = TC.fwd_bwd(TC.reference_chain, x, theta, go, in_range, dt)
            out[f"{name}/{tag}/out"] = r["out"].numpy()
            out[f"{name}/{tag}/grad_x"] = r["grad_x"].numpy()
            if theta is not None:
                out[f"{name}/{tag}/grad_theta"] = r["grad_theta"].numpy()
        names.append(name)

    # known-answer micro-vectors, all through the real reference functions
    red = torch.tensor([1., 0., 0., 1.]).view(4, 1, 1).expand(4, 2, 2)
    green = torch.tensor([0., 1., 0., 1.]).view(4, 1, 1).expand(4, 2, 2)
    out["ka/order/in"] = torch.stack([red, green])[None].numpy()            # layer 0 = back
    out["ka/order/out"] = iu.alpha_composite_pytorch(torch.stack([red, green])[None].clone()).numpy()
    clear = torch.zeros(1, 3, 4, 2, 2)
    clear[:, :, :3] = 0.7
    out["ka/transparent/in"] = clear.numpy()
    out["ka/transparent/out"] = iu.alpha_composite_pytorch(clear.clone()).numpy()
    half = torch.tensor([[1., 0., 0., .5], [0., 1., 0., .5], [0., 0., 1., .5]]).view(3, 4, 1, 1).expand(3, 4, 2, 2)
    out["ka/half/in"] = half[None].numpy()                                   # image_utils.py:413-420
    out["ka/half/out"] = iu.alpha_composite_pytorch(half[None].clone()).numpy()
    tr = torch.tensor([[[0.5, -0.25], [0.0, 1.0]]])
    out["ka/translate2x3/in"] = tr.numpy()
    out["ka/translate2x3/out"] = iu.convert_translate_to_2x3(tr).numpy()
    # +tx moves content left (image_utils.py:23-28): an impulse at column 5 of a [0,1] image
    imp = torch.zeros(1, 1, 4, 8, 8)
    imp[0, 0, :, 4, 5] = 1.0
    th = iu.convert_translate_to_2x3(torch.tensor([[[0.5, 0.0]]]))
    grid = torch.nn.functional.affine_grid(th.view(-1, 2, 3), (1, 4, 8, 8), align_corners=False)
    out["ka/shift/in"] = imp.numpy()
    out["ka/shift/theta"] = th.numpy()
    out["ka/shift/out"] = torch.nn.functional.grid_sample(imp.view(1, 4, 8, 8), grid, align_corners=False).numpy()

    out["cases"] = np.array(names)
    out["torch_version"] = np.array(torch.__version__)
    dst = os.path.join(ROOT, "tests", "golden", "render_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
