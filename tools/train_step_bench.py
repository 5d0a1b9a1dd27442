#!/usr/bin/env python
"""BASELINE config 4: the global-GAN training step -- placement net (STNv2c, this library's warp) + renderer +
global discriminator + Adam -- under DistributedDataParallel, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_bench.py

One step = a G phase (non-saturating loss through D and the renderer into the placement net; the placement net's
16 MB of fp32 gradients are all-reduced) followed by a D phase (fake stack detached + a real stack composited without a
warp; D's 126 MB of gradients are all-reduced), as the reference alternates them (custom/training_loop_aio.py:446-492,
custom/loss_aio.py:280-341 without the lazy R1 term).  The collective is DDP's NCCL all-reduce over NVLink; the renderer
itself never communicates.  Reported per wiring of the renderer:
    fused    STNv2c(fused=True) -> FusedRenderer(x, theta)          warp + composite in one kernel
    swap     STNv2c -> AnalyticRenderer                              materialised warp, then composite (class swap only)
    aten     the reference's chain on the same tensors (ATen affine_grid + grid_sample + batched a_over_b)
and, for the fused wiring, the step again under DDP's no_sync(): the difference is what the all-reduces cost
(`allreduce_share`).

The discriminator has the reference's aio architecture and size -- custom.networks_aio.Discriminator(img_resolution=256,
img_channels=4, init_res=[8, 8], channel_base=16384, channel_max=512) (train_aio.py:179, 209-215): residual blocks
256 -> 16 with channels 64-128-256-512-512-512, minibatch-stddev epilogue at 8 x 8, 31.6 M parameters = 126 MB of fp32
gradients -- written with plain torch.nn convolutions: the reference builds it on its bias_act / upfirdn2d plugins, which
are out of scope and absent on the GPU box; the FIR resampling of its down-sampling convolutions is
replaced by stride 2, which changes neither the parameter count nor the all-reduce.  Random layer stacks stand in for
the nine local generators.  Adam betas (0, 0.99) as in train_aio.py:217-220, but lr 1e-5: with the reference's 0.0025 the
freshly initialised placement net throws every layer off the canvas within two steps (translations of +-150), after
which there is nothing left to render or differentiate."""
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


class DBlock(nn.Module):
    """One residual block of the StyleGAN2 discriminator (custom/networks_aio.py: DiscriminatorBlock, 'resnet')."""

    def __init__(self, cin, cout, first):
        super().__init__()
        self.fromrgb = nn.Conv2d(4, cin, 1) if first else None
        self.conv0 = nn.Conv2d(cin, cin, 3, padding=1)
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=2, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1, stride=2, bias=False)

    def forward(self, x):
        if self.fromrgb is not None:
            x = F.leaky_relu(self.fromrgb(x), 0.2)
        y = self.skip(x) * (0.5 ** 0.5)
        x = F.leaky_relu(self.conv1(F.leaky_relu(self.conv0(x), 0.2)), 0.2) * (0.5 ** 0.5)
        return x + y


class GlobalD(nn.Module):
    """The aio global discriminator's shape: 256 -> 8 in five residual blocks, minibatch stddev, conv, two linears."""

    def __init__(self):
        super().__init__()
        ch = {256: 64, 128: 128, 64: 256, 32: 512, 16: 512, 8: 512}
        res = [256, 128, 64, 32, 16]
        self.blocks = nn.Sequential(*[DBlock(ch[r], ch[r // 2], r == 256) for r in res])
        self.conv = nn.Conv2d(ch[8] + 1, ch[8], 3, padding=1)
        self.fc = nn.Linear(ch[8] * 8 * 8, ch[8])
        self.out = nn.Linear(ch[8], 1)

    def forward(self, img):
        x = self.blocks(img)
        g = min(4, x.shape[0])
        y = x.reshape(g, -1, *x.shape[1:])
        y = (y - y.mean(0)).square().mean(0).add(1e-8).sqrt().mean([1, 2, 3]).reshape(-1, 1, 1, 1)
        x = torch.cat([x, y.repeat(g, 1, x.shape[2], x.shape[3])], 1)
        x = F.leaky_relu(self.conv(x), 0.2)
        return self.out(F.leaky_relu(self.fc(x.flatten(1)), 0.2))


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
    import contextlib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L, R = 16, 9, 256                                     # per GPU (batch_gpu of the reference's aio config is 4..16)
    torch.manual_seed(rank)
    x = synth.make_layers(4, L, R, R, "F", seed=rank).repeat(B // 4, 1, 1, 1, 1).to(dev)
    real = synth.make_layers(4, L, R, R, "F", seed=100 + rank).repeat(B // 4, 1, 1, 1, 1).to(dev)
    res = {"n_gpus": world, "per_gpu_batch": B, "layers": L, "resolution": R}
    for mode in ("fused", "swap", "aten"):
        stn = M.STNv2c(R, 4, L, fused=(mode != "swap")).to(dev)
        with torch.no_grad():                                # leave the identity placement so that theta gets a gradient signal
            stn.fc_loc[2].bias.normal_(0, 0.1)
        D = GlobalD().to(dev)
        n_stn, n_d = sum(p.numel() for p in stn.parameters()), sum(p.numel() for p in D.parameters())
        renderer = M.FusedRenderer(R, 4, L) if mode == "fused" else M.AnalyticRenderer(R, 4, L)
        compositor = M.AnalyticRenderer(R, 4, L)             # the real branch: layers already in place, no warp
        if world > 1:
            stn = nn.parallel.DistributedDataParallel(stn, device_ids=[local])
            D = nn.parallel.DistributedDataParallel(D, device_ids=[local])
        opt_g = torch.optim.Adam(stn.parameters(), lr=1e-5, betas=(0.0, 0.99))
        opt_d = torch.optim.Adam(D.parameters(), lr=1e-5, betas=(0.0, 0.99))

        def render(y, theta):
            return renderer(y, theta) if mode == "fused" else (renderer(y) if mode == "swap" else aten_chain(y, theta))

        def step(sync=True):
            ctx = contextlib.nullcontext if (sync or world == 1) else None
            with (contextlib.ExitStack() if ctx is None else ctx()) as stack:
                if ctx is None:
                    stack.enter_context(stn.no_sync())
                    stack.enter_context(D.no_sync())
                # G phase (loss_aio.py:285-289): D's weights frozen, gradients reach the placement net through the renderer
                # (D is called as a plain module here: DDP would wait for gradients that this phase does not produce)
                Dm = D.module if world > 1 else D
                Dm.requires_grad_(False)
                opt_g.zero_grad(set_to_none=True)
                y, theta = stn(x)
                loss_g = F.softplus(-Dm(render(y, theta))).mean()
                loss_g.backward()
                opt_g.step()
                Dm.requires_grad_(True)
                # D phase (loss_aio.py:299-320): fake stack detached, real stack composited without a warp
                opt_d.zero_grad(set_to_none=True)
                with torch.no_grad():
                    y, theta = stn(x)
                    fake = render(y, theta)
                loss_d = F.softplus(D(fake)).mean() + F.softplus(-D(compositor(real))).mean()
                loss_d.backward()
                opt_d.step()
            return loss_g

        def timed(sync, n=8):
            for _ in range(3):
                step(sync)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                loss = step(sync).detach()
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms), float(loss)

        ms, loss = timed(True)
        inner = stn.module if world > 1 else stn
        gnorm = float(sum(p.grad.float().norm() ** 2 for p in inner.parameters() if p.grad is not None) ** 0.5)
        res[mode] = {"ms_per_step": round(ms, 3), "images_per_s": round(world * B / ms * 1e3, 1), "loss_g": round(loss, 5),
                     "placement_grad_norm": float(f"{gnorm:.3e}")}
        if mode == "fused":
            res["params"] = {"placement_net": n_stn, "global_D": n_d, "allreduce_MB_per_step": round((n_stn + n_d) * 4 / 1e6, 1)}
            if world > 1:
                ms_ns, _ = timed(False)
                res[mode].update(ms_per_step_no_allreduce=round(ms_ns, 3), allreduce_share=round(max(0.0, 1 - ms_ns / ms), 3))
        del stn, D, opt_g, opt_d
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
