"""AugmentPipe's geometric execution block on the GPU.

``geometric_warp(images, G_inv)`` is the part of ``training.augment.AugmentPipe.forward`` that turns the accumulated
inverse transform into pixels (``training/augment.py:306-342``): reflect pad, x2 upsample through the sym6 low-pass,
affine bilinear resampling, low-pass + x2 decimation + crop.  The reference does it with ``F.pad``, two ``upfirdn2d``
calls and ``affine_grid`` + ``grid_sample``; here it is three kernels forward and four backward behind
``mgr_augment_geom_*``.  The host-side arithmetic below (padding margins, the matrix handed to the sampler) follows the
reference line by line -- like the reference it reads the margins back to the host, one small synchronisation.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

HZ_PAD = 3        # len(sym6) // 4  (augment.py:316)


def _translate(tx, ty):
    return torch.tensor([[1.0, 0.0, tx], [0.0, 1.0, ty], [0.0, 0.0, 1.0]], dtype=torch.float32)


def _scale(sx, sy):
    return torch.tensor([[sx, 0.0, 0.0], [0.0, sy, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)


def padding_margins(G_inv: torch.Tensor, H: int, W: int):
    """(mx0, my0, mx1, my1) of augment.py:311-322: where G_inv sends the image corners, the farthest excursion over the
    batch plus twice the filter reach, clamped to [0, size - 1], rounded up."""
    cx, cy = (W - 1) / 2, (H - 1) / 2
    cp = torch.tensor([[-cx, -cy, 1.0], [cx, -cy, 1.0], [cx, cy, 1.0], [-cx, cy, 1.0]], dtype=torch.float32)
    cp = G_inv @ cp.t()
    m = cp[:, :2, :].permute(1, 0, 2).flatten(1)
    m = torch.cat([-m, m]).max(dim=1).values
    m = m + torch.tensor([HZ_PAD * 2 - cx, HZ_PAD * 2 - cy] * 2, dtype=torch.float32)
    m = m.max(torch.zeros(4)).min(torch.tensor([W - 1, H - 1] * 2, dtype=torch.float32))
    return tuple(int(v) for v in m.ceil().to(torch.int32))


def sampling_theta(G_inv: torch.Tensor, H: int, W: int, mx0: int, my0: int, mx1: int, my1: int) -> torch.Tensor:
    """[B,2,3] for the sampler (augment.py:326-338): G_inv re-centred on the padded image, moved to the x2 grid and to
    the normalised coordinates of the upsampled input and of the 2(H+6) x 2(W+6) output grid."""
    G = _translate((mx0 - mx1) / 2, (my0 - my1) / 2) @ G_inv
    G = _scale(2, 2) @ G @ _scale(0.5, 0.5)
    G = _translate(-0.5, -0.5) @ G @ _translate(0.5, 0.5)
    Hu, Wu = 2 * (H + my0 + my1), 2 * (W + mx0 + mx1)
    Hs, Ws = 2 * (H + 2 * HZ_PAD), 2 * (W + 2 * HZ_PAD)
    G = _scale(2 / Wu, 2 / Hu) @ G @ _scale(Ws / 2, Hs / 2)
    return G[:, :2, :].contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _launch(entry: str, x: torch.Tensor, theta: torch.Tensor, margins) -> torch.Tensor:
    """One call of ``mgr_augment_geom_forward`` / ``_backward`` on a detached, contiguous fp32 ``x [B,C,H,W]``."""
    lib = _lib.load()
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    nbytes = lib.mgr_augment_geom_workspace_bytes(B, C, H, W, *margins)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = getattr(lib, entry)(_p(x), _p(theta), _p(y), _p(ws), nbytes, B, C, H, W, *margins,
                                 ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc, entry)
    _lib.launch_count += 1
    return y


# The block is LINEAR in the images (pad, FIR, bilinear resampling, FIR): its backward is the adjoint operator applied
# to grad_out, and the backward of THAT is the forward operator applied to the incoming cotangent.  The two Functions
# below call each other, so the block differentiates to any order -- what the reference gets from
# grid_sample_gradfix (torch_utils/ops/grid_sample_gradfix.py:49-88) and upfirdn2d's own double backward, and what the
# R1 penalty needs: autograd.grad(real_logits.sum(), real_layer_tmp, create_graph=True) runs through the augment pipe
# that sits between the renderer and global D (custom/loss_aio.py:252-254, 327-338).
class _GeometricWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, theta, margins):
        ctx.margins = margins
        ctx.save_for_backward(theta)
        return _launch("mgr_augment_geom_forward", images.detach().contiguous(), theta, margins)

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.needs_input_grad[0]:
            return None, None, None
        (theta,) = ctx.saved_tensors
        return _GeometricWarpAdjoint.apply(grad_out, theta, ctx.margins), None, None


class _GeometricWarpAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_out, theta, margins):
        ctx.margins = margins
        ctx.save_for_backward(theta)
        return _launch("mgr_augment_geom_backward", grad_out.detach().to(torch.float32).contiguous(), theta, margins)

    @staticmethod
    def backward(ctx, grad_grad):
        if not ctx.needs_input_grad[0]:
            return None, None, None
        (theta,) = ctx.saved_tensors
        return _GeometricWarp.apply(grad_grad.to(torch.float32), theta, ctx.margins), None, None


def geometric_warp(images: torch.Tensor, G_inv: torch.Tensor) -> torch.Tensor:
    """``images [B,C,H,W]`` fp32 on the GPU, ``G_inv [B,3,3]`` (or ``[3,3]``): the inverse pixel-space transform the pipe
    has accumulated.  Returns the transformed images; differentiable w.r.t. ``images`` to any order (G_inv is random, not learned).
    A G_inv that is the identity object short-circuits in the reference (augment.py:309); pass it anyway and the block
    runs as a mild low-pass, or skip the call as the reference does."""
    if images.dim() != 4 or images.dtype != torch.float32:
        raise ValueError("images must be a float32 [B,C,H,W] tensor")
    if not images.is_cuda:
        raise _lib.MontageRenderError("images must be a CUDA tensor: no CPU path")
    B, C, H, W = images.shape
    G = G_inv.detach().to("cpu", torch.float32)
    if G.dim() == 2:
        G = G.expand(B, 3, 3)
    if tuple(G.shape) != (B, 3, 3):
        raise ValueError(f"G_inv must be [B,3,3] = {(B, 3, 3)}, got {tuple(G_inv.shape)}")
    margins = padding_margins(G, H, W)
    theta = sampling_theta(G, H, W, *margins).to(images.device)
    return _GeometricWarp.apply(images, theta, margins)
