#!/usr/bin/env python
"""Throughput of mgr_composite_u8 (the 8-bit Pillow-exact composite) against its HBM roofline (developer tool)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import render as mr, synth  # noqa: E402
from bench import measured_peak_gbs  # noqa: E402

peak, _ = measured_peak_gbs()
for (B, L, H, W, dt) in ((64, 7, 256, 256, torch.bfloat16), (64, 7, 256, 256, torch.float32), (32, 16, 512, 512, torch.float32)):
    xs = [synth.make_layers(8, L, H, W, "S", seed=s).repeat(B // 8, 1, 1, 1, 1).to("cuda", dt) for s in range(3)]
    for x in xs:
        mr.alpha_composite(x, in_range="m11")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for k in range(n):
        mr.alpha_composite(xs[k % 3], in_range="m11")
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / n
    nbytes = B * L * 4 * H * W * xs[0].element_size() + B * 4 * H * W * 4
    print(json.dumps({"wl": f"B{B} L{L} {H}x{W} {str(dt).replace('torch.', '')}", "us": round(us, 1), "GBs": round(nbytes / us / 1e3, 1),
                      "frac_of_measured_peak": round(nbytes / us / 1e3 / peak, 3), "layer_Mpix_s": round(B * L * H * W / us)}))
