#!/usr/bin/env python
"""Timeline of the chunked host pipeline (developer tool): replays mgr_render_fwd_bwd_host's schedule with torch
streams and timing events and prints, per chunk, when the H2D / kernels / D2H started and ended (ms from the start)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import synth  # noqa: E402
from montage_gan_b200.render import _Render  # noqa: E402

B, L, H, W = 64, 7, 256, 256
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nslot = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dt = torch.bfloat16
x = synth.make_layers(8, L, H, W, "S", seed=0).repeat(8, 1, 1, 1, 1).to(dt).pin_memory()
th = synth.make_theta(B, L, "I", seed=0).pin_memory()
go = synth.make_grad_out(B, H, W, seed=0).to(dt).pin_memory()
h_out = torch.empty((B, 4, H, W), dtype=dt).pin_memory()
h_gx = torch.empty((B, L, 4, H, W), dtype=dt).pin_memory()
h_gt = torch.empty((B, L, 2, 3)).pin_memory()
dev = torch.device("cuda:0")
slots = [dict(x=torch.empty((chunk, L, 4, H, W), dtype=dt, device=dev), th=torch.empty((chunk, L, 2, 3), device=dev),
              go=torch.empty((chunk, 4, H, W), dtype=dt, device=dev)) for _ in range(nslot)]
s_in, s_out, s_c = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def step(record):
    n = B // chunk
    ev = [[E() for _ in range(6)] for _ in range(n)]
    comp = [None] * nslot
    freed = [None] * nslot
    res = [None] * nslot
    t0 = E(); t0.record()
    for s in (s_in, s_out, s_c):
        s.wait_stream(torch.cuda.current_stream())
    for c in range(n):
        k, b0 = c % nslot, c * chunk
        S = slots[k]
        with torch.cuda.stream(s_in):
            if comp[k] is not None:
                s_in.wait_event(comp[k])
            ev[c][0].record()
            S["x"].copy_(x[b0:b0 + chunk], non_blocking=True)
            S["th"].copy_(th[b0:b0 + chunk], non_blocking=True)
            S["go"].copy_(go[b0:b0 + chunk], non_blocking=True)
            ev[c][1].record()
        with torch.cuda.stream(s_c):
            s_c.wait_event(ev[c][1])
            if freed[k] is not None:
                s_c.wait_event(freed[k])
            ev[c][2].record()
            xs = S["x"].detach().requires_grad_(True); ts = S["th"].detach().requires_grad_(True)
            out = montage_gan_b200.render.render(xs, ts)
            gx, gt = torch.autograd.grad(out, (xs, ts), S["go"])
            ev[c][3].record()
            comp[k] = ev[c][3]
            res[k] = (out, gx, gt)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev[c][3])
            ev[c][4].record()
            h_out[b0:b0 + chunk].copy_(out.detach(), non_blocking=True)
            h_gx[b0:b0 + chunk].copy_(gx, non_blocking=True)
            h_gt[b0:b0 + chunk].copy_(gt, non_blocking=True)
            ev[c][5].record()
            freed[k] = ev[c][5]
    torch.cuda.current_stream().wait_stream(s_out)
    t1 = E(); t1.record()
    torch.cuda.synchronize()
    if record:
        for c in range(n):
            print(json.dumps({"chunk": c, **{nm: round(t0.elapsed_time(e), 3) for nm, e in zip(("h2d0", "h2d1", "k0", "k1", "d2h0", "d2h1"), ev[c])}}))
    return t0.elapsed_time(t1)


for _ in range(3):
    step(False)
print("total ms", round(step(True), 3))
