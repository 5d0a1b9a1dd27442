#!/usr/bin/env python
"""Ceiling for the e2e figure: pinned-host <-> device copy bandwidth on this box, each direction alone and both at
once, in one piece and in pipeline-sized chunks, with and without the render kernels running beside the copies
(developer tool; the numbers go into DESIGN.md next to the e2e line)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, piece=n, iters=6, kernels=None):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in (s1, s2, s3):
        s.wait_stream(torch.cuda.current_stream())
    for _ in range(iters):
        for off in range(0, n, piece):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[off:off + piece].copy_(h_in[off:off + piece], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[off:off + piece].copy_(d_out[off:off + piece], non_blocking=True)
            if kernels is not None:
                with torch.cuda.stream(s3):
                    kernels()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return round(n * iters / 1e6 / e0.elapsed_time(e1), 1)   # GB/s per direction


def make_kernels():
    import montage_gan_b200  # noqa: F401
    from montage_gan_b200 import synth
    from montage_gan_b200.render import render
    x = synth.make_layers(8, 7, 256, 256, "S", seed=0).to(torch.bfloat16).cuda().requires_grad_(True)
    th = synth.make_theta(8, 7, "I", seed=0).cuda().requires_grad_(True)
    go = synth.make_grad_out(8, 256, 256, seed=0).to(torch.bfloat16).cuda()

    def k():
        out = render(x, th)
        torch.autograd.grad(out, (x, th), go)
    return k


run(True, True, n, 2)
res = {"h2d_alone": run(True, False), "d2h_alone": run(False, True), "both": run(True, True)}
for mb in (8, 16, 32, 64):
    res[f"both_{mb}MB_pieces"] = run(True, True, mb << 20)
k = make_kernels()
k(); torch.cuda.synchronize()
res["both_32MB_pieces_with_kernels"] = run(True, True, 32 << 20, kernels=k)
res["h2d_32MB_pieces_with_kernels"] = run(True, False, 32 << 20, kernels=k)
print(json.dumps(res))
