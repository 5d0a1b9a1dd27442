#!/usr/bin/env python
"""Developer stress run: the kernels on TMA box copies against the staged ones (mgr_set_debug_path(4)) on random shapes,
dtypes, range modes and shift magnitudes, fused renderer and materialised warp.  Prints the worst deviations."""
import sys, os, random
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa
from montage_gan_b200 import _lib, render as mr, synth

lib = _lib.load()
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60


def rel(a, b):
    s = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / s


worst = {}
for it in range(N):
    B, L = rng.randint(1, 5), rng.choice([2, 3, 5, 7, 9, 12, 16, 20, 24, 32])
    H = rng.choice([8, 15, 16, 17, 31, 33, 48, 64, 100, 130, 257])
    W = 4 * rng.randint(2, 80)
    dt = rng.choice([torch.float32, torch.bfloat16, torch.float16])
    mode = rng.choice(["m11", "01"])
    scale = rng.choice([0.05, 0.5, 1.0, 2.5])
    x = synth.make_layers(B, L, H, W, "S", seed=it)
    if mode == "01":
        x = (x + 1) / 2
    g = torch.Generator().manual_seed(it)
    th = torch.eye(2, 3).expand(B, L, 2, 3).clone()
    th[..., 2] = (torch.rand(B, L, 2, generator=g) * 2 - 1) * scale
    th[:, 0, :, 2] = torch.tensor([0.31 * 2 / W, -0.27 * 2 / H])
    go = synth.make_grad_out(B, H, W, seed=it)
    gw = torch.randn(B, L, 4, H, W, generator=g)
    res = {}
    for path in (0, 4):
        _lib.check(lib.mgr_set_debug_path(path), "path")
        xd = x.cuda().to(dt).requires_grad_(True); td = th.cuda().requires_grad_(True)
        out = mr.render(xd, td, in_range=mode); out.backward(go.cuda().to(dt))
        xw = x.cuda().to(dt).requires_grad_(True); tw = th.cuda().requires_grad_(True)
        w = mr.warp(xw, tw, in_range=mode); w.backward(gw.cuda().to(dt))
        torch.cuda.synchronize()
        res[path] = [t.detach().float().cpu().numpy() for t in (out, xd.grad, td.grad, w, xw.grad, tw.grad)]
    lib.mgr_set_debug_path(0)
    names = ["out", "grad_x", "grad_theta", "warp", "warp_grad_x", "warp_grad_theta"]
    for n_, a, b in zip(names, res[0], res[4]):
        if not np.isfinite(a).all():
            print("NON-FINITE", n_, (B, L, H, W), dt, mode, scale); continue
        e = float(np.abs(a - b).max()) if n_ in ("out", "warp") else rel(a, b)
        key = (n_, str(dt).replace("torch.", ""))
        if e > worst.get(key, (0,))[0]:
            worst[key] = (e, (B, L, H, W), mode, scale)
for k in sorted(worst):
    print(k, worst[k])
