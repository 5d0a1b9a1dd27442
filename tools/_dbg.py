import sys, os, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import montage_gan_b200
from montage_gan_b200 import synth, render as mr
def run(x, th, go, need_t):
    xr = x.detach().requires_grad_(True); tr = th.detach().requires_grad_(need_t)
    out = mr.render(xr, tr)
    g = torch.autograd.grad(out, (xr, tr) if need_t else (xr,), go)
    return out.detach(), g[0]
for (B, L, H, W, dt, fam) in ((2, 9, 256, 256, torch.float32, "F"), (2, 5, 256, 256, torch.float32, "F"), (2, 9, 256, 256, torch.float32, "S"),
                              (2, 9, 128, 128, torch.float32, "F"), (1, 9, 256, 256, torch.float32, "F")):
    x = synth.make_layers(B, L, H, W, fam, seed=5).cuda().to(dt)
    th = synth.make_theta(B, L, "T", seed=5, cover_back=False).cuda()
    go = synth.make_grad_out(B, H, W, "randn", seed=5).cuda().to(dt)
    for need_t in (True, False):
        o0, g0 = run(x, th, go, need_t)
        bad = 0; chans = set(); badout = 0
        for it in range(6):
            o1, g1 = run(x, th, go, need_t)
            d = (g1 != g0)
            bad += int(d.sum()); badout += int((o1 != o0).sum())
            if d.any(): chans |= set(torch.nonzero(d)[:, 2].tolist())
        print(f"B{B} L{L} {H}x{W} {fam} need_theta={need_t}: differing grad_x elements over 6 reruns = {bad} (channels {sorted(chans)}), out diffs {badout}")
