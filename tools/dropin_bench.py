#!/usr/bin/env python
"""The two ways a maintainer can wire the library in, fwd+bwd at C2 size:
  (a) drop-in classes only: STNv2c returns the MATERIALISED warped layers (mgr_warp_*), AnalyticRenderer composites them
  (b) fused: STNv2c(fused=True) hands (x, theta) to FusedRenderer -- the warped layers never exist
(the localisation CNN is left out of both: thetas are given).  Developer tool -> DESIGN.md."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import render as mr, synth  # noqa: E402


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


B, L, H, W = 64, 7, 256, 256
PATH = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # mgr_set_debug_path: 4 = without the kernels on TMA box copies
from montage_gan_b200 import _lib  # noqa: E402
_lib.check(_lib.load().mgr_set_debug_path(PATH), 'mgr_set_debug_path')
for dt in (torch.bfloat16, torch.float32):
    for tf in ("T", "I"):
        x = synth.make_layers(8, L, H, W, "S", seed=0).repeat(8, 1, 1, 1, 1).to("cuda", dt)
        th = synth.make_theta(B, L, tf, seed=0, cover_back=False).cuda()
        go = synth.make_grad_out(B, H, W, seed=0).to("cuda", dt)

        def two_step():
            xr, tr = x.detach().requires_grad_(True), th.detach().requires_grad_(True)
            out = mr.render(mr.warp(xr, tr), None)
            torch.autograd.grad(out, (xr, tr), go)

        def fused():
            xr, tr = x.detach().requires_grad_(True), th.detach().requires_grad_(True)
            out = mr.render(xr, tr)
            torch.autograd.grad(out, (xr, tr), go)

        def warp_fwd_only():
            mr.warp(x, th)

        a, b, c = timeit(two_step), timeit(fused), timeit(warp_fwd_only)
        print(json.dumps({"path": PATH, "dtype": str(dt).replace("torch.", ""), "theta": tf, "warp_then_composite_us": round(a, 1), "fused_us": round(b, 1),
                          "warp_forward_only_us": round(c, 1)}), flush=True)
