#!/usr/bin/env python
""""also time the incumbent GPU path" -- the reference's own chain (ATen affine_grid + grid_sample,
then the per-sample / per-layer a_over_b loop of custom_utils/image_utils.py:128-163, autograd backward) with the
tensors on the B200, i.e. the existing Blackwell kernels this library replaces.  Developer measurement, not a bench
arm: prints one JSON line per workload, CUDA-event timed, for profiles/ and DESIGN.md.

The chain is written out here (a dozen lines of torch) rather than imported from oracle/, which stays reserved for
the tests and the CPU arm of bench.py.  It also runs the batched form (one a_over_b per layer over the whole batch --
what a maintainer would write first) so the comparison is not only against Python-loop overhead."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import synth  # noqa: E402
from montage_gan_b200.render import render  # noqa: E402


def a_over_b(t1, t2, ch):                                   # image_utils.py:128-133 (ch = channel axis)
    c1, a1 = t1.narrow(ch, 0, 3), t1.narrow(ch, 3, 1)
    c2, a2 = t2.narrow(ch, 0, 3), t2.narrow(ch, 3, 1)
    ao = a1 + a2 * (1 - a1)
    return torch.cat([torch.nan_to_num((c1 * a1 + c2 * a2 * (1 - a1)) / ao), ao], ch)


def warp(x, theta):                                          # fukuwarai/networks.py:247-258
    B, L, C, H, W = x.shape
    x2 = x.reshape(-1, C, H, W)
    grid = F.affine_grid(theta.reshape(-1, 2, 3), x2.size(), align_corners=False)
    return (F.grid_sample(x2 + 1, grid, align_corners=False) - 1).view(B, L, C, H, W)


def chain_reference_loops(x, theta):                         # image_utils.py:142-146, :163 + loss_aio.py:251
    z = (warp(x, theta) + 1.) / 2.
    outs = []
    for lchw in z:
        canvas = lchw[0]
        for chw in lchw[1:]:
            canvas = a_over_b(chw, canvas, 0)
        outs.append(canvas)
    return torch.stack(outs) * 2. - 1.


def chain_batched(x, theta):
    z = (warp(x, theta) + 1.) / 2.
    canvas = z[:, 0]
    for l in range(1, z.shape[1]):
        canvas = a_over_b(z[:, l], canvas, 1)
    return canvas * 2. - 1.


def time_fn(fn, iters):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    for (B, L, H, W, dt, name) in ((8, 7, 256, 256, torch.float32, "C1"), (64, 7, 256, 256, torch.float32, "C2 sizes, fp32 as the reference computes"),
                                   (64, 7, 256, 256, torch.bfloat16, "C2 bf16 storage")):
        x = synth.make_layers(min(B, 8), L, H, W, "S", seed=0).repeat(B // min(B, 8), 1, 1, 1, 1).to(dev, dt)
        th = synth.make_theta(B, L, "I", seed=0).to(dev)
        go = synth.make_grad_out(B, H, W, seed=0).to(dev, dt)
        res = {"workload": name, "B": B, "L": L, "H": H, "W": W, "dtype": str(dt).replace("torch.", "")}
        mpix = B * L * H * W / 1e6

        def run(chain, xx, tt, gg):
            xr, tr = xx.detach().requires_grad_(True), tt.detach().requires_grad_(True)
            out = chain(xr, tr)
            torch.autograd.grad(out, (xr, tr), gg)

        if dt == torch.float32:
            for nm, chain in (("aten_reference_loops", chain_reference_loops), ("aten_batched", chain_batched)):
                ms = time_fn(lambda: run(chain, x, th, go), 3)
                res[nm] = {"ms": round(ms, 3), "layer_Mpix_s": round(mpix / 1e-3 / ms / 1e0, 1)}
        ms = time_fn(lambda: run(render, x, th, go), 20)
        res["this_library"] = {"ms": round(ms, 3), "layer_Mpix_s": round(mpix / 1e-3 / ms, 1)}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
