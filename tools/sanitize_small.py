"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import montage_gan_b200
from montage_gan_b200 import render as mr, synth, _lib
lib = _lib.load()
dev = "cuda:0"
for dtype in (torch.float32, torch.bfloat16):
    for (B, L, H, W) in ((2, 3, 40, 36), (1, 7, 64, 64), (2, 2, 33, 20)):
        for fam in ("I", "P", "X"):
            x = synth.make_layers(B, L, H, W, "F", seed=1).to(dev, dtype).requires_grad_(True)
            th = (synth.make_theta(B, L, "T", seed=1, cover_back=False) if fam == "P" else synth.make_theta(B, L, fam, seed=1)).to(dev).requires_grad_(True)
            for path in (0, 1, 2):
                lib.mgr_set_debug_path(path)
                out = mr.render(x, th)
                out.backward(torch.randn_like(out))
            lib.mgr_set_debug_path(0)
            w = mr.warp(x, th); w.sum().backward()
        xs = synth.make_layers(B, L, H, W, "S", seed=2).to(dev, dtype).requires_grad_(True)
        go = torch.zeros(B, 4, H, W, device=dev, dtype=dtype, requires_grad=True)
        o = mr.render(xs, None)
        (gx,) = torch.autograd.grad(o, xs, go, create_graph=True)
        gx.square().sum().backward()
mr.make_batch_for_pos_estimator([torch.rand(2, 4, 20, 12, device=dev), torch.rand(2, 4, 32, 32, device=dev)], -1, canvas=(32, 32))
torch.cuda.synchronize()
print("sanitize_small: done")
