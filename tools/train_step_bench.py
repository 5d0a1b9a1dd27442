#!/usr/bin/env python
"""SURVEY.md 8(d) config 4 in miniature: one global-G style training step -- placement net (STNv2c, this library's
warp) + renderer + a discriminator + Adam -- under DistributedDataParallel, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py

What it shows: gradients flow from the discriminator's logits through the renderer's custom autograd Function into the
placement net, DDP's NCCL all-reduce of the placement net's gradients works around it, and what the renderer costs
inside a step, for the three ways of wiring it:
    fused    STNv2c(fused=True) -> FusedRenderer(x, theta)          warp + composite in one kernel
    swap     STNv2c -> AnalyticRenderer                              materialised warp, then composite (class swap only)
    aten     the reference's chain on the same tensors (ATen affine_grid + grid_sample + batched a_over_b)
The discriminator is a STAND-IN (a small strided-conv net): the reference's global D is a StyleGAN2 discriminator built
on its own bias_act / upfirdn2d plugins, which are out of scope and absent on the GPU box.
Random layer stacks stand in for the nine local generators.  Adam betas (0, 0.99) as in train_aio.py:217-220, but lr 1e-5:
with the reference's 0.0025 the freshly initialised placement net throws every layer off the canvas within two steps
(translations of +-150), after which there is nothing left to render or differentiate."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import montage_gan_b200  # noqa: F401,E402
from montage_gan_b200 import modules as M, synth  # noqa: E402


class StandInD(nn.Module):
    def __init__(self, ch=64):
        super().__init__()
        layers, c = [], 4
        for k in range(6):                                   # 256 -> 4
            layers += [nn.Conv2d(c, min(ch * 2 ** k, 512), 4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
            c = min(ch * 2 ** k, 512)
        self.body = nn.Sequential(*layers)
        self.head = nn.Linear(c * 4 * 4, 1)

    def forward(self, img):
        return self.head(self.body(img).flatten(1))


def aten_chain(x, theta):
    B, L, C, H, W = x.shape
    x2 = x.reshape(-1, C, H, W)
    grid = F.affine_grid(theta.reshape(-1, 2, 3), x2.size(), align_corners=False)
    z = ((F.grid_sample(x2 + 1, grid, align_corners=False) - 1).view(B, L, C, H, W) + 1) / 2
    canvas = z[:, 0]
    for l in range(1, L):
        c1, a1, c2, a2 = z[:, l, :3], z[:, l, 3:], canvas[:, :3], canvas[:, 3:]
        ao = a1 + a2 * (1 - a1)
        canvas = torch.cat([torch.nan_to_num((c1 * a1 + c2 * a2 * (1 - a1)) / ao), ao], 1)
    return canvas * 2 - 1


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl")
    B, L, R = 16, 9, 256                                     # per GPU (batch_gpu of the reference's aio config is 4..16)
    torch.manual_seed(rank)
    x = synth.make_layers(4, L, R, R, "F", seed=rank).repeat(B // 4, 1, 1, 1, 1).to(dev)
    res = {"n_gpus": world, "per_gpu_batch": B, "layers": L, "resolution": R}
    for mode in ("fused", "swap", "aten"):
        stn = M.STNv2c(R, 4, L, fused=(mode != "swap")).to(dev)
        with torch.no_grad():                                # leave the identity placement so that theta gets a gradient signal
            stn.fc_loc[2].bias.normal_(0, 0.1)
        D = StandInD().to(dev)
        renderer = M.FusedRenderer(R, 4, L) if mode == "fused" else M.AnalyticRenderer(R, 4, L)
        if world > 1:
            stn = nn.parallel.DistributedDataParallel(stn, device_ids=[local])
        opt = torch.optim.Adam(stn.parameters(), lr=1e-5, betas=(0.0, 0.99))

        def step():
            opt.zero_grad(set_to_none=True)
            y, theta = stn(x)
            img = renderer(y, theta) if mode == "fused" else (renderer(y) if mode == "swap" else aten_chain(y, theta))
            loss = F.softplus(-D(img)).mean()                # non-saturating G loss (loss_aio.py:285-289)
            loss.backward()                                  # (detach nothing: D's weights also get grads, as in Gmain)
            opt.step()
            return loss

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            loss = step().detach()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        inner = stn.module if world > 1 else stn
        gnorm = float(sum(p.grad.float().norm() ** 2 for p in inner.parameters() if p.grad is not None) ** 0.5)
        res[mode] = {"ms_per_step": round(float(ms), 3), "images_per_s": round(world * B / float(ms) * 1e3, 1),
                     "loss": round(float(loss), 5), "placement_grad_norm": float(f"{gnorm:.3e}")}
        del stn, D, opt
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
