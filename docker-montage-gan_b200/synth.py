"""Synthetic layer stacks and placements (SURVEY.md 8d "Synthetic inputs").

Deterministic (CPU ``torch.Generator``) so the oracle and the CUDA path see identical bits.
Layer families follow what the reference feeds the path: ``x`` is the tanh-range output of the
local generators padded with -1 (``custom_utils/image_utils.py:229-243``); thetas follow
``image_utils.random_position`` (``:281-294``, translation U(-1,1)) and the "random affine
params" of BASELINE.json config 1.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# zoom-in placement for layer 0 so every output pixel is covered (finite reference gradients,
# SURVEY.md 8c "NaN-free oracle configurations")
COVER_THETA = ((0.8, 0.05, 0.03), (-0.05, 0.8, -0.02))

# alpha>0 area fractions of the reference's 9 fixture layers (custom_utils/test_data/layers)
FIXTURE_FRACTIONS = (0.44, 0.18, 0.02, 0.25, 0.025, 0.010, 0.002, 0.18, 0.005)


def _gen(seed):
    return torch.Generator().manual_seed(int(seed))


def _smooth(t, k=15):
    pad = k // 2
    shape = t.shape
    t = t.reshape(-1, 1, shape[-2], shape[-1])
    ker = torch.ones(1, 1, k, k) / (k * k)
    for _ in range(2):
        t = F.conv2d(F.pad(t, (pad, pad, pad, pad), mode="reflect"), ker)
    t = t.reshape(shape)
    lo = t.amin(dim=(-2, -1), keepdim=True)
    hi = t.amax(dim=(-2, -1), keepdim=True)
    return (t - lo) / (hi - lo).clamp_min(1e-12)


def make_layers(B, L, H, W, family="W", seed=0, alpha_min=0.05):
    """[B,L,4,H,W] float32 in [-1,1].  family: 'W' white noise, 'S' smooth, 'F' fixture-like
    sparse masks (alpha exactly 0 outside a blob, exactly 1 inside, soft edge)."""
    g = _gen(seed)
    if family == "W":
        z = torch.rand(B, L, 4, H, W, generator=g)
        z[:, :, 3] = alpha_min + (1 - alpha_min) * z[:, :, 3]
    elif family == "S":
        k = max(3, (min(H, W) // 16) * 2 + 1)
        z = _smooth(torch.rand(B, L, 4, H, W, generator=g), k=min(k, 15))
        z[:, :, 3] = alpha_min + (1 - alpha_min) * z[:, :, 3]
    elif family == "F":
        z = torch.rand(B, L, 4, H, W, generator=g)
        yy = torch.linspace(-1, 1, H).view(1, 1, H, 1)
        xx = torch.linspace(-1, 1, W).view(1, 1, 1, W)
        frac = torch.tensor([FIXTURE_FRACTIONS[l % len(FIXTURE_FRACTIONS)] for l in range(L)])
        rad = (frac * 4 / math.pi).sqrt().view(1, L, 1, 1)
        cx = (torch.rand(B, L, 1, 1, generator=g) - 0.5) * 0.6
        cy = (torch.rand(B, L, 1, 1, generator=g) - 0.5) * 0.6
        d = ((xx - cx) ** 2 + (yy - cy) ** 2).sqrt()
        edge = 4.0 / min(H, W)
        z[:, :, 3] = ((rad - d) / edge + 0.5).clamp(0, 1)
    else:
        raise ValueError(family)
    return (z * 2 - 1).contiguous()


def make_theta(B, L, family="I", seed=0, cover_back=True):
    """[B,L,2,3] float32.  'I' = I + 0.25 N(0,1); 'T' translation U(-1,1); '0' identity;
    'X' extreme: scale 2^U(-2,2), rotation U(0,2pi), shift U(-1,1)."""
    g = _gen(seed + 1000)
    eye = torch.eye(2, 3).expand(B, L, 2, 3).clone()
    if family == "I":
        th = eye + 0.25 * torch.randn(B, L, 2, 3, generator=g)
    elif family == "T":
        th = eye
        th[..., 2] = torch.rand(B, L, 2, generator=g) * 2 - 1
    elif family == "0":
        th = eye
    elif family == "X":
        s = 2 ** (torch.rand(B, L, generator=g) * 4 - 2)
        r = torch.rand(B, L, generator=g) * 2 * math.pi
        th = torch.zeros(B, L, 2, 3)
        th[..., 0, 0] = s * r.cos()
        th[..., 0, 1] = -s * r.sin()
        th[..., 1, 0] = s * r.sin()
        th[..., 1, 1] = s * r.cos()
        th[..., 2] = torch.rand(B, L, 2, generator=g) * 2 - 1
    else:
        raise ValueError(family)
    if cover_back and family != "0":
        th[:, 0] = torch.tensor(COVER_THETA)
    return th.contiguous()


def make_grad_out(B, H, W, kind="randn", seed=0):
    if kind == "ones":
        return torch.ones(B, 4, H, W)
    return torch.randn(B, 4, H, W, generator=_gen(seed + 2000))
