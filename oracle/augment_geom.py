"""CPU restatement of AugmentPipe's geometric execution block (TEST INFRASTRUCTURE ONLY).

Reference: ``training/augment.py:306-342`` -- given a batch ``images [B,C,H,W]`` and the inverse pixel-space transform
``G_inv [B,3,3]`` the pipe has accumulated (flip / rotate90 / integer + fractional translation / scale / rotation /
anisotropy), it
  1. sizes a reflect padding from where G_inv sends the image corners (``:311-322``),
  2. reflect-pads and upsamples x2 with the 12-tap ``sym6`` low-pass (``upfirdn2d.upsample2d``, ``:325-331``),
  3. resamples with ``affine_grid`` + bilinear ``grid_sample`` onto a ``2(H+6) x 2(W+6)`` grid (``:333-340``),
  4. low-pass filters, decimates x2 and crops back to ``H x W`` (``upfirdn2d.downsample2d``, ``:342``).
The FIR arithmetic lives in ``torch_utils/ops/upfirdn2d.py:168-222`` (``_upfirdn2d_ref``: zero insertion, pad / crop,
two 1-D convolutions, decimation); restated below with plain convolutions.  Pinned by
``tests/golden/augment_geom_golden.npz`` (``oracle/make_golden_augment.py``: the reference pipe itself, run on CPU).

``geometric_warp(..., double_backward=True)`` swaps ATen's ``grid_sample`` (whose backward has no derivative -- the
reason the reference carries ``torch_utils/ops/grid_sample_gradfix.py``) for ``bilinear_sample`` below, the same
arithmetic written with index gathers, so that the R1 pattern (``custom/loss_aio.py:327-338``) can be differentiated
twice on the CPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# Daubechies least-asymmetric wavelet sym6, decomposition low-pass (PyWavelets ``pywt.Wavelet('sym6').dec_lo``); the
# reference keeps the same table at training/augment.py:50-52 and normalises it to unit DC gain (upfirdn2d.py:121-122).
SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148]


def lowpass_filter() -> torch.Tensor:
    f = torch.as_tensor(SYM6, dtype=torch.float32)
    return f / f.sum()


def _conv_sep(x, f):
    """``x [B,C,H,W]`` correlated with the 1-D taps ``f`` along W then along H, 'valid' (upfirdn2d.py:211-216)."""
    C = x.shape[1]
    w = f.to(x.dtype)[None, None].repeat(C, 1, 1)
    x = F.conv2d(x, w.unsqueeze(2), groups=C)
    return F.conv2d(x, w.unsqueeze(3), groups=C)


def upsample2x(x, f):
    """``upfirdn2d.upsample2d(x, f, up=2)`` (upfirdn2d.py:327-362): zeros between samples, pad (fw+1)//2 = 6 before and
    (fw-2)//2 = 5 after, convolve (= correlate with the flipped taps), gain up^2 split over the two 1-D passes."""
    B, C, H, W = x.shape
    u = x.new_zeros(B, C, 2 * H, 2 * W)
    u[:, :, ::2, ::2] = x
    u = F.pad(u, [6, 5, 6, 5])
    return _conv_sep(u, f.flip(0) * 2.0)


def downsample2x(x, f, pad):
    """``upfirdn2d.downsample2d(x, f, down=2, padding=pad, flip_filter=True)`` (upfirdn2d.py:365-400): pad (crop when
    negative) by (fw-1)//2 + pad = 5 + pad on both sides, correlate with the taps as given, keep every second sample."""
    p = (len(f) - 2 + 1) // 2 + pad
    q = (len(f) - 2) // 2 + pad
    assert p <= 0 and q <= 0, "the block only ever crops here"
    x = x[:, :, -p: x.shape[2] + q, -p: x.shape[3] + q]
    return _conv_sep(x, f)[:, :, ::2, ::2]


def _translate(tx, ty):
    return torch.tensor([[1.0, 0.0, tx], [0.0, 1.0, ty], [0.0, 0.0, 1.0]], dtype=torch.float32)


def _scale(sx, sy):
    return torch.tensor([[sx, 0.0, 0.0], [0.0, sy, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)


def margins(G_inv: torch.Tensor, H: int, W: int, hz_pad: int = 3):
    """Reflect padding (mx0, my0, mx1, my1) the batch needs (augment.py:311-322): the image corners through G_inv, the
    farthest excursion over the batch plus the filter reach, clamped to [0, size - 1], rounded up."""
    cx, cy = (W - 1) / 2, (H - 1) / 2
    cp = torch.tensor([[-cx, -cy, 1.0], [cx, -cy, 1.0], [cx, cy, 1.0], [-cx, cy, 1.0]], dtype=torch.float32)
    cp = G_inv @ cp.t()                                              # [B, xyz, idx]
    m = cp[:, :2, :].permute(1, 0, 2).flatten(1)                      # [xy, B * idx]
    m = torch.cat([-m, m]).max(dim=1).values                          # [x0, y0, x1, y1]
    m = m + torch.tensor([hz_pad * 2 - cx, hz_pad * 2 - cy] * 2, dtype=torch.float32)
    m = m.max(torch.zeros(4)).min(torch.tensor([W - 1, H - 1] * 2, dtype=torch.float32))
    return tuple(int(v) for v in m.ceil().to(torch.int32))


def sampling_theta(G_inv, H, W, mx0, my0, mx1, my1, hz_pad: int = 3):
    """The 2x3 matrices handed to affine_grid (augment.py:326-338): G_inv moved to the padded image's centre, to the
    x2 grid (pixel centres shift by half a pixel), and to normalised coordinates of the input / output grids."""
    G = _translate((mx0 - mx1) / 2, (my0 - my1) / 2) @ G_inv
    G = _scale(2, 2) @ G @ _scale(0.5, 0.5)
    G = _translate(-0.5, -0.5) @ G @ _translate(0.5, 0.5)
    Hu, Wu = 2 * (H + my0 + my1), 2 * (W + mx0 + mx1)                 # upsampled padded image
    Hs, Ws = 2 * (H + 2 * hz_pad), 2 * (W + 2 * hz_pad)               # resampling grid
    G = _scale(2 / Wu, 2 / Hu) @ G @ _scale(Ws / 2, Hs / 2)
    return G[:, :2, :], (Hs, Ws)


def bilinear_sample(x: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """``F.grid_sample(x, grid, 'bilinear', 'zeros', align_corners=False)`` with plain indexing (ATen
    ``GridSampler.h:27-36, 205-207``: unnormalise, four corners, out-of-range corners contribute 0).  Linear in ``x`` and
    built from ops that autograd differentiates to any order."""
    B, C, H, W = x.shape
    ix = ((grid[..., 0] + 1) * W - 1) / 2
    iy = ((grid[..., 1] + 1) * H - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    fx, fy = (ix - x0)[:, None], (iy - y0)[:, None]
    x0, y0 = x0.long(), y0.long()
    flat = x.reshape(B, C, H * W)

    def tap(yy, xx):
        inside = ((xx >= 0) & (xx < W) & (yy >= 0) & (yy < H))[:, None].to(x.dtype)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).reshape(B, 1, -1).expand(B, C, -1)
        return torch.gather(flat, 2, idx).reshape(B, C, *grid.shape[1:3]) * inside

    return (tap(y0, x0) * (1 - fx) * (1 - fy) + tap(y0, x0 + 1) * fx * (1 - fy) +
            tap(y0 + 1, x0) * (1 - fx) * fy + tap(y0 + 1, x0 + 1) * fx * fy)


def geometric_warp(images: torch.Tensor, G_inv: torch.Tensor, hz_pad: int = 3, double_backward: bool = False) -> torch.Tensor:
    """The whole block for ``images [B,C,H,W]`` (fp32 / fp64, CPU) and ``G_inv [B,3,3]``; differentiable via autograd
    (twice with ``double_backward=True``)."""
    B, C, H, W = images.shape
    f = lowpass_filter()
    mx0, my0, mx1, my1 = margins(G_inv.float(), H, W, hz_pad)
    x = F.pad(images, [mx0, mx1, my0, my1], mode="reflect")
    x = upsample2x(x, f)
    theta, (Hs, Ws) = sampling_theta(G_inv.float(), H, W, mx0, my0, mx1, my1, hz_pad)
    grid = F.affine_grid(theta.to(images.dtype), [B, C, Hs, Ws], align_corners=False)
    if double_backward:
        x = bilinear_sample(x, grid)
    else:
        x = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return downsample2x(x, f, -hz_pad * 2)


def recover_G_inv(theta, H, W, mx0, my0, mx1, my1, hz_pad: int = 3):
    """Inverse of ``sampling_theta`` (used by the golden generator, which can only observe what the reference hands
    to affine_grid)."""
    B = theta.shape[0]
    G = torch.cat([theta.float(), torch.tensor([[[0.0, 0.0, 1.0]]]).expand(B, 1, 3)], 1)
    Hu, Wu = 2 * (H + my0 + my1), 2 * (W + mx0 + mx1)
    Hs, Ws = 2 * (H + 2 * hz_pad), 2 * (W + 2 * hz_pad)
    G = _scale(Wu / 2, Hu / 2) @ G @ _scale(2 / Ws, 2 / Hs)
    G = _translate(0.5, 0.5) @ G @ _translate(-0.5, -0.5)
    G = _scale(0.5, 0.5) @ G @ _scale(2, 2)
    return _translate(-(mx0 - mx1) / 2, -(my0 - my1) / 2) @ G
