#!/usr/bin/env python
"""AugmentPipe's geometric block (mgr_augment_geom_*) against the same chain written with ATen ops on the same GPU
(F.pad reflect, zero-insertion + two grouped 1-D convolutions per FIR stage as in upfirdn2d's reference implementation,
affine_grid + grid_sample) -- the reference's own upfirdn2d CUDA plugin is not part of this repo.  fwd+bwd w.r.t. the
images, CUDA-event timed (developer tool -> DESIGN.md)."""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import augment as A  # noqa: E402

SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148]


def conv_sep(x, f):
    C = x.shape[1]
    w = f[None, None].repeat(C, 1, 1)
    return F.conv2d(F.conv2d(x, w.unsqueeze(2), groups=C), w.unsqueeze(3), groups=C)


def aten_chain(x, theta, margins, f):
    mx0, my0, mx1, my1 = margins
    B, C, H, W = x.shape
    x = F.pad(x, [mx0, mx1, my0, my1], mode="reflect")
    u = x.new_zeros(B, C, 2 * x.shape[2], 2 * x.shape[3])
    u[:, :, ::2, ::2] = x
    u = conv_sep(F.pad(u, [6, 5, 6, 5]), f.flip(0) * 2)
    grid = F.affine_grid(theta, [B, C, 2 * (H + 6), 2 * (W + 6)], align_corners=False)
    s = F.grid_sample(u, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return conv_sep(s[:, :, 1:-1, 1:-1], f)[:, :, ::2, ::2]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


f = torch.tensor(SYM6, device="cuda")
f = f / f.sum()
for (B, C, H, W) in ((64, 4, 256, 256), (16, 4, 256, 256)):
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(B, C, H, W, generator=g) * 2 - 1).cuda()
    ang = (torch.rand(B, generator=g) - 0.5) * 1.0
    Gi = torch.eye(3).repeat(B, 1, 1)
    Gi[:, 0, 0], Gi[:, 0, 1], Gi[:, 1, 0], Gi[:, 1, 1] = torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang)
    Gi[:, 0, 2] = (torch.rand(B, generator=g) - 0.5) * 0.25 * W
    go = torch.randn(B, C, H, W, generator=g).cuda()
    margins = A.padding_margins(Gi, H, W)
    theta = A.sampling_theta(Gi, H, W, *margins).cuda()

    def ours():
        xr = x.detach().requires_grad_(True)
        torch.autograd.grad(A._GeometricWarp.apply(xr, theta, margins), xr, go)

    def aten():
        xr = x.detach().requires_grad_(True)
        torch.autograd.grad(aten_chain(xr, theta, margins, f), xr, go)

    def ours_fwd():
        A._GeometricWarp.apply(x, theta, margins)

    a, b, c = timeit(ours), timeit(aten), timeit(ours_fwd)
    mx0, my0, mx1, my1 = margins
    up = B * C * 4 * (H + my0 + my1) * (W + mx0 + mx1)
    fwd_bytes = 4 * (B * C * H * W * 2 + 2 * up + 2 * B * C * 4 * (H + 6) * (W + 6))     # images in/out + U and S written and read once
    print(json.dumps({"B": B, "C": C, "H": H, "W": W, "margins": margins, "ours_fwd_bwd_us": round(a, 1), "aten_fwd_bwd_us": round(b, 1),
                      "speedup": round(b / a, 2), "ours_fwd_us": round(c, 1), "fwd_GBs_incl_intermediates": round(fwd_bytes / c / 1e3)}), flush=True)
