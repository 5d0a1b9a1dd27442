#!/usr/bin/env python
"""Ragged stack (the reference's nine face-part sizes on a 256 x 256 canvas, custom/dataset_aio.py:28-83) against the
padded-canvas path: fwd+bwd through the autograd bindings, CUDA-event timed (developer tool -> DESIGN.md)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import render as mr, synth  # noqa: E402

SIZES = [(256, 256), (256, 256), (160, 224), (256, 256), (96, 160), (64, 96), (64, 32), (256, 256), (64, 160)]
H = W = 256
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


for dt in (torch.bfloat16, torch.float32):
    for tf in ("T", "I"):
        full = synth.make_layers(8, len(SIZES), H, W, "S", seed=1).repeat(B // 8, 1, 1, 1, 1)
        layers = [full[:, l, :, (H - h) // 2:(H - h) // 2 + h, (W - w) // 2:(W - w) // 2 + w].contiguous().to("cuda", dt)
                  for l, (h, w) in enumerate(SIZES)]
        theta = synth.make_theta(B, len(SIZES), tf, seed=1, cover_back=False).cuda()
        go = synth.make_grad_out(B, H, W, "randn", seed=1).to("cuda", dt)

        def ragged():
            xs = [t.detach().requires_grad_(True) for t in layers]
            th = theta.detach().requires_grad_(True)
            out = mr.render_ragged(xs, th, canvas=(H, W))
            torch.autograd.grad(out, xs + [th], go)

        def canvas(pad_inside=True):
            xs = [t.detach().requires_grad_(True) for t in layers]
            th = theta.detach().requires_grad_(True)
            x = mr.make_batch_for_pos_estimator(xs, pad_value=-1, canvas=(H, W))
            out = mr.render(x, th)
            torch.autograd.grad(out, xs + [th], go)

        padded = mr.make_batch_for_pos_estimator(layers, pad_value=-1, canvas=(H, W)).detach()

        def canvas_only():
            x = padded.detach().requires_grad_(True)
            th = theta.detach().requires_grad_(True)
            out = mr.render(x, th)
            torch.autograd.grad(out, (x, th), go)

        r, c, co = timeit(ragged), timeit(canvas), timeit(canvas_only)
        px = B * sum(h * w for h, w in SIZES)
        print(json.dumps({"dtype": str(dt).replace("torch.", ""), "theta": tf, "B": B, "ragged_us": round(r, 1),
                          "pad_then_canvas_us": round(c, 1), "canvas_render_only_us": round(co, 1),
                          "speedup_vs_pad_then_canvas": round(c / r, 2), "native_layer_Mpix_s": round(px / r),
                          "canvas_layer_Mpix_s": round(B * len(SIZES) * H * W / co)}), flush=True)
