"""CPU restatement of the reference's analytic render path (warp + alpha-over), numpy only.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it, and only as the checker / the CPU arm.  The product path
(``docker-montage-gan_b200``) never imports this module and has no CPU fallback.

What is restated (paths relative to /root/reference/montage_gan):

* the warp of every layer by its 2x3 placement, ``fukuwarai/networks.py:247-258`` (STNv2c, the
  ``+1 / grid_sample / -1`` form) and ``:217-226`` (STNv2b, plain form) -- whose arithmetic
  lives in a third-party dependency that is NOT under /root/reference: PyTorch ATen
  ``affine_grid_generator`` + ``grid_sampler_2d`` (+ ``_backward``).  The reference pins
  PyTorch 1.8.0a0 (``Dockerfile:9``, nvcr.io/nvidia/pytorch:20.12-py3); this container has
  torch 2.11.0.  The published algorithm restated here: base grid
  ``linspace(-1,1,n)*(n-1)/n``; ``grid = base @ theta^T``; un-normalise
  ``ix = ((gx+1)*W-1)/2`` (``ATen/native/GridSampler.h:27-36``); bilinear with zeros padding
  (``GridSampler.h:205-207``).
* straight-alpha "over" compositing, layer 0 = back, ``custom_utils/image_utils.py:112-163``
  (``a_over_b`` 128-133, ``process`` 142-146, batched 163), and the range shifts
  ``image_utils.py:184-195`` used at ``custom/loss_aio.py:251``.
* theta builders ``image_utils.py:316-335`` (``convert_translate_to_2x3``).

Backward is the analytic adjoint, written independently of autograd,
and pinned against ``torch.autograd`` through the real reference in ``tests/`` (in this
container, where /root/reference exists) and through the committed golden vectors under
``tests/golden/`` (generated from the real reference by ``oracle/make_golden.py``).

Pinning status: the reference has NO tests or golden vectors for this path;
parity is pinned by outputs of the reference itself run in the build container.

Everything is dtype-generic: ``dtype=np.float32`` follows the reference's fp32 op order
(bitwise for the grid, <=2e-6 for the sampler on this CPU), ``dtype=np.float64`` is "truth".

Deliberate, documented deviation (SURVEY.md finding 3): where the composited alpha is exactly
0 the reference's backward is NaN (0/0 inside ``a_over_b``); this restatement -- like the CUDA
kernels -- defines every gradient as 0 there.  ``nan_mask`` reports those pixels so tests can
assert the reference is non-finite exactly there.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "base_coords", "affine_grid", "grid_sample_fwd", "grid_sample_bwd", "affine_grid_bwd",
    "a_over_b", "alpha_composite", "alpha_composite_bwd", "normalize_minus11",
    "normalize_zero1", "convert_translate_to_2x3", "warp_fwd", "render_fwd",
    "render_fwd_bwd", "composite_jvp",
]


def _t(dtype):
    return np.dtype(dtype).type


# --------------------------------------------------------------------------------------
# warp: ATen affine_grid_generator + grid_sampler_2d (call sites fukuwarai/networks.py:251-254)
# --------------------------------------------------------------------------------------

def base_coords(n: int, dtype=np.float32) -> np.ndarray:
    """ATen ``linspace_from_neg_one(n, align_corners=False)`` = ``linspace(-1,1,n)*(n-1)/n``.

    ``torch.linspace`` fills from both ends (idx < n/2: ``start+step*idx``, else
    ``end-step*(n-1-idx)``) with a fused multiply-add; reproduced bitwise in fp32
    (checked for n in 1..1024 in tests/test_oracle.py).  Mathematically ``(2j+1)/n - 1``.
    """
    dt = _t(dtype)
    if n == 1:
        return np.zeros(1, dt)
    step = dt(dt(2) / dt(n - 1))
    idx = np.arange(n)
    wide = np.longdouble if dt is np.float64 else np.float64  # emulate single rounding (FMA)
    lo = (wide(-1) + wide(step) * idx).astype(dt)
    hi = (wide(1) - wide(step) * (n - 1 - idx)).astype(dt)
    lin = np.where(idx < n // 2, lo, hi).astype(dt)
    return ((lin * dt(n - 1)).astype(dt) / dt(n)).astype(dt)


def affine_grid(theta: np.ndarray, H: int, W: int, dtype=np.float32):
    """``F.affine_grid(theta[N,2,3], (N,C,H,W), align_corners=False)`` -> (gx, gy) each [N,H,W].

    The CPU bmm with K=3 accumulates ``x*t0``, then fma(y,t1,.), then ``+t2`` (measured
    bitwise in this container); fp64 mode is insensitive to the order.
    """
    dt = _t(dtype)
    wide = np.longdouble if dt is np.float64 else np.float64
    th = np.asarray(theta, dtype=dt).reshape(-1, 2, 3)
    x = base_coords(W, dt)[None, None, :]
    y = base_coords(H, dt)[None, :, None]
    out = []
    for k in range(2):
        t0 = th[:, k, 0][:, None, None]
        t1 = th[:, k, 1][:, None, None]
        t2 = th[:, k, 2][:, None, None]
        acc = (x * t0).astype(dt)
        acc = (wide(1) * y * t1 + acc).astype(dt)          # fma(y, t1, acc)
        acc = (acc + t2).astype(dt)
        out.append(np.broadcast_to(acc, (th.shape[0], H, W)).copy())
    return out[0], out[1]


def _unnormalize(g, size, dt):
    # ATen grid_sampler_unnormalize, align_corners=False: ((coord + 1) * size - 1) / 2
    return ((((g + dt(1)).astype(dt) * dt(size)).astype(dt) - dt(1)).astype(dt) / dt(2)).astype(dt)


def _corner_aux(gx, gy, H, W, dt):
    ix = _unnormalize(gx, W, dt)
    iy = _unnormalize(gy, H, dt)
    x0f = np.floor(ix)
    y0f = np.floor(iy)
    fx = (ix - x0f).astype(dt)
    fy = (iy - y0f).astype(dt)
    # clip before the integer cast so absurd thetas (config-5 stress) cannot overflow int64
    x0 = np.clip(x0f, -2, W + 1).astype(np.int64)
    y0 = np.clip(y0f, -2, H + 1).astype(np.int64)
    return x0, y0, fx, fy


def _gather(img, yy, xx):
    """img [N,C,H,W]; yy,xx [N,H,W] int -> [N,C,H,W], zero where out of bounds."""
    N, C, H, W = img.shape
    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    lin = (np.clip(yy, 0, H - 1) * W + np.clip(xx, 0, W - 1)).reshape(N, 1, -1)
    v = np.take_along_axis(img.reshape(N, C, H * W), np.broadcast_to(lin, (N, C, lin.shape[2])), axis=2)
    v = v.reshape(N, C, yy.shape[1], yy.shape[2])
    return np.where(ok[:, None], v, img.dtype.type(0)), ok


def grid_sample_fwd(img: np.ndarray, gx: np.ndarray, gy: np.ndarray):
    """``F.grid_sample(img, grid, mode='bilinear', padding_mode='zeros', align_corners=False)``.

    img [N,C,H,W], gx/gy [N,Ho,Wo].  Returns (out [N,C,Ho,Wo], aux) where aux carries what
    the backward needs.  Accumulation order nw, ne, sw, se with fused multiply-adds (matches
    the ATen CPU vector kernel on 99.1 % of outputs bitwise, max 1.8e-6, measured here).
    """
    dt = img.dtype.type
    wide = np.longdouble if dt is np.float64 else np.float64
    N, C, H, W = img.shape
    x0, y0, fx, fy = _corner_aux(gx, gy, H, W, dt)
    ex = (dt(1) - fx).astype(dt)
    ey = (dt(1) - fy).astype(dt)
    nw, ok_nw = _gather(img, y0, x0)
    ne, ok_ne = _gather(img, y0, x0 + 1)
    sw, ok_sw = _gather(img, y0 + 1, x0)
    se, ok_se = _gather(img, y0 + 1, x0 + 1)
    w_nw = (ex * ey).astype(dt)[:, None]
    w_ne = (fx * ey).astype(dt)[:, None]
    w_sw = (ex * fy).astype(dt)[:, None]
    w_se = (fx * fy).astype(dt)[:, None]
    o = (nw * w_nw).astype(dt)
    for v, w in ((ne, w_ne), (sw, w_sw), (se, w_se)):
        o = (wide(1) * v * w + o).astype(dt)
    aux = dict(x0=x0, y0=y0, fx=fx, fy=fy, corners=(nw, ne, sw, se),
               ok=(ok_nw, ok_ne, ok_sw, ok_se), weights=(w_nw, w_ne, w_sw, w_se))
    return o, aux


def grid_sample_bwd(img_shape, aux, gout: np.ndarray):
    """Adjoint of ``grid_sample_fwd`` (ATen ``grid_sampler_2d_backward``).

    Returns (grad_img [N,C,H,W], ggx [N,Ho,Wo], ggy [N,Ho,Wo]).  Scatter accumulates in
    float64 (summation order of the reference's sequential/atomic adds is not contractual).
    """
    dt = gout.dtype.type
    N, C, H, W = img_shape
    x0, y0, fx, fy = aux["x0"], aux["y0"], aux["fx"], aux["fy"]
    nw, ne, sw, se = aux["corners"]          # already zero where out of bounds
    gimg = np.zeros(N * C * H * W, np.float64)
    plane = (np.arange(N)[:, None] * C + np.arange(C)[None, :])[:, :, None] * (H * W)   # [N,C,1]
    for (dy, dx), w, ok in zip(((0, 0), (0, 1), (1, 0), (1, 1)), aux["weights"], aux["ok"]):
        yy = np.clip(y0 + dy, 0, H - 1)
        xx = np.clip(x0 + dx, 0, W - 1)
        lin = (yy * W + xx).reshape(N, 1, -1) + plane                                   # [N,C,HoWo]
        contrib = (gout * w * ok[:, None]).reshape(N, C, -1)
        gimg += np.bincount(lin.ravel(), weights=contrib.ravel().astype(np.float64), minlength=gimg.size)
    ex = (dt(1) - fx)[:, None]
    ey = (dt(1) - fy)[:, None]
    fxe = fx[:, None]
    fye = fy[:, None]
    dix = (gout * ((ne - nw) * ey + (se - sw) * fye)).sum(axis=1, dtype=np.float64)
    diy = (gout * ((sw - nw) * ex + (se - ne) * fxe)).sum(axis=1, dtype=np.float64)
    ggx = (dix * (W / 2.0)).astype(dt)
    ggy = (diy * (H / 2.0)).astype(dt)
    return gimg.reshape(N, C, H, W).astype(dt), ggx, ggy


def affine_grid_bwd(ggx: np.ndarray, ggy: np.ndarray, dtype=np.float32) -> np.ndarray:
    """Adjoint of ``affine_grid``: grad_theta[N,2,3] = sum_ij [gg * x_j, gg * y_i, gg]."""
    dt = _t(dtype)
    N, H, W = ggx.shape
    x = base_coords(W, dt).astype(np.float64)[None, None, :]
    y = base_coords(H, dt).astype(np.float64)[None, :, None]
    gt = np.zeros((N, 2, 3), np.float64)
    for k, gg in enumerate((ggx.astype(np.float64), ggy.astype(np.float64))):
        gt[:, k, 0] = (gg * x).sum(axis=(1, 2))
        gt[:, k, 1] = (gg * y).sum(axis=(1, 2))
        gt[:, k, 2] = gg.sum(axis=(1, 2))
    return gt.astype(dt)


# --------------------------------------------------------------------------------------
# composite: custom_utils/image_utils.py:112-163 (default, non-premultiplied branch)
# --------------------------------------------------------------------------------------

def _nan_to_num(v):
    # torch.nan_to_num defaults: nan -> 0, +-inf -> +-finfo.max  (image_utils.py:100-109)
    return np.nan_to_num(v, nan=0.0, posinf=np.finfo(v.dtype).max, neginf=np.finfo(v.dtype).min)


def a_over_b(chw1: np.ndarray, chw2: np.ndarray) -> np.ndarray:
    """One "over" step, ``image_utils.py:128-133``; arrays [...,4,H,W], chw1 in front of chw2."""
    dt = chw1.dtype.type
    c1, a1 = chw1[..., :3, :, :], chw1[..., 3:, :, :]
    c2, a2 = chw2[..., :3, :, :], chw2[..., 3:, :, :]
    one_m = (dt(1) - a1).astype(dt)
    alpha_out = (a1 + (a2 * one_m).astype(dt)).astype(dt)
    num = ((c1 * a1).astype(dt) + ((c2 * a2).astype(dt) * one_m).astype(dt)).astype(dt)
    with np.errstate(divide="ignore", invalid="ignore"):
        color_out = _nan_to_num((num / alpha_out).astype(dt))
    return np.concatenate([color_out, alpha_out], axis=-3)


def alpha_composite(blchw: np.ndarray) -> np.ndarray:
    """``alpha_composite_pytorch(blchw)`` (iterative, straight alpha, layer 0 = back),
    ``image_utils.py:142-146, 163``.  [B,L,4,H,W] (or [L,4,H,W]) in [0,1] -> [B,4,H,W]."""
    z = np.asarray(blchw)
    unbatched = z.ndim == 4
    if unbatched:
        z = z[None]
    canvas = z[:, 0]
    for l in range(1, z.shape[1]):
        canvas = a_over_b(z[:, l], canvas)
    return canvas[0] if unbatched else canvas


def alpha_composite_bwd(blchw: np.ndarray, gout: np.ndarray):
    """Analytic adjoint of ``alpha_composite``, division-free in (1-a_l).

    Returns (grad_blchw, nan_mask[B,H,W]) -- nan_mask marks pixels whose composited alpha is
    exactly 0 (reference backward is NaN there; we define 0).  For L == 1 the reference
    returns the layer untouched, so the adjoint is the identity.
    """
    z = np.asarray(blchw)
    dt = z.dtype.type
    B, L = z.shape[:2]
    if L == 1:
        return gout[:, None].astype(dt).copy(), np.zeros((B,) + z.shape[-2:], bool)
    c = z[:, :, :3].astype(np.float64)
    a = z[:, :, 3:].astype(np.float64)
    S = np.zeros((B, L + 1, 3) + z.shape[-2:])
    R = np.zeros((B, L + 1, 1) + z.shape[-2:])
    for l in range(L):
        S[:, l + 1] = a[:, l] * c[:, l] + (1 - a[:, l]) * S[:, l]
        R[:, l + 1] = a[:, l] + (1 - a[:, l]) * R[:, l]
    T = np.ones((B, L, 1) + z.shape[-2:])
    for l in range(L - 2, -1, -1):
        T[:, l] = T[:, l + 1] * (1 - a[:, l + 1])
    A = R[:, L]
    P = S[:, L]
    zero = A == 0
    Asafe = np.where(zero, 1.0, A)
    o = P / Asafe
    g = gout.astype(np.float64)
    GP = np.where(zero, 0.0, g[:, :3] / Asafe)
    GA = np.where(zero, 0.0, g[:, 3:] - (g[:, :3] * o).sum(axis=1, keepdims=True) / Asafe)
    gz = np.zeros(z.shape, np.float64)
    for l in range(L):
        gz[:, l, :3] = GP * T[:, l] * a[:, l]
        gz[:, l, 3:] = T[:, l] * ((GP * (c[:, l] - S[:, l])).sum(axis=1, keepdims=True) + GA * (1 - R[:, l]))
    return gz.astype(dt), zero[:, 0]


def composite_jvp(blchw: np.ndarray, tangent: np.ndarray) -> np.ndarray:
    """Forward-mode derivative of ``alpha_composite`` along ``tangent`` (same shape as blchw).

    Used to check the double-backward (R1 penalty on the real branch,
    ``custom/loss_aio.py:313-341``): the backward is linear in grad_out, so its adjoint with
    respect to grad_out is this JVP.  Computed in float64 by differentiating the closed form.
    """
    z = np.asarray(blchw, np.float64)
    dz = np.asarray(tangent, np.float64)
    B, L = z.shape[:2]
    if L == 1:
        return dz[:, 0].astype(blchw.dtype)
    S = np.zeros((B, 3) + z.shape[-2:]); dS = np.zeros_like(S)
    R = np.zeros((B, 1) + z.shape[-2:]); dR = np.zeros_like(R)
    for l in range(L):
        c, a = z[:, l, :3], z[:, l, 3:]
        dc, da = dz[:, l, :3], dz[:, l, 3:]
        dS = da * c + a * dc - da * S + (1 - a) * dS
        S = a * c + (1 - a) * S
        dR = da - da * R + (1 - a) * dR
        R = a + (1 - a) * R
    zero = R == 0
    Rs = np.where(zero, 1.0, R)
    do_rgb = np.where(zero, 0.0, dS / Rs - S * dR / (Rs * Rs))
    return np.concatenate([do_rgb, dR], axis=1).astype(blchw.dtype)


# --------------------------------------------------------------------------------------
# range shifts and theta builders
# --------------------------------------------------------------------------------------

def normalize_minus11(t):
    """``image_utils.py:184-188``: [0,1] -> [-1,1]."""
    dt = t.dtype.type
    return ((t * dt(2)).astype(t.dtype) - dt(1)).astype(t.dtype)


def normalize_zero1(t):
    """``image_utils.py:191-195``: [-1,1] -> [0,1]."""
    dt = t.dtype.type
    return ((t + dt(1)).astype(t.dtype) / dt(2)).astype(t.dtype)


def convert_translate_to_2x3(translation: np.ndarray) -> np.ndarray:
    """``image_utils.py:316-335``: [...,2] (dx, dy) -> [...,2,3] = [[1,0,dx],[0,1,dy]]."""
    tr = np.asarray(translation)
    theta = np.zeros(tr.shape[:-1] + (2, 3), tr.dtype)
    theta[..., 0, 0] = 1
    theta[..., 1, 1] = 1
    theta[..., :, 2] += tr
    return theta


# --------------------------------------------------------------------------------------
# the chained hot path
# --------------------------------------------------------------------------------------

def warp_fwd(x: np.ndarray, theta: np.ndarray, in_range: str = "m11", dtype=np.float32):
    """The warp half only: ``fukuwarai/networks.py:250-257`` (m11, STNv2c) or ``:219-225``
    ('01', STNv2b / ``image_utils.random_position`` 281-294).  Returns (warped, aux)."""
    dt = _t(dtype)
    x = np.asarray(x, dt)
    B, L, C, H, W = x.shape
    gx, gy = affine_grid(np.asarray(theta, dt).reshape(B * L, 2, 3), H, W, dt)
    x2 = x.reshape(B * L, C, H, W)
    if in_range == "m11":
        s2, aux = grid_sample_fwd((x2 + dt(1)).astype(dt), gx, gy)
        w = (s2 - dt(1)).astype(dt)
    elif in_range == "01":
        w, aux = grid_sample_fwd(x2, gx, gy)
    else:
        raise ValueError(in_range)
    return w.reshape(B, L, C, H, W), aux


def render_fwd(x, theta=None, in_range: str = "m11", dtype=np.float32) -> np.ndarray:
    """Warp (if theta is given) then composite, the chain of SURVEY.md 3.2:
    ``STNv2c.forward`` warp (``fukuwarai/networks.py:250-257``) followed by
    ``normalize_minus11(alpha_composite_pytorch(normalize_zero1(.)))`` (``custom/loss_aio.py:251``).
    ``theta=None`` is the real-image branch (composite only, ``loss_aio.py:313-320``)."""
    dt = _t(dtype)
    x = np.asarray(x, dt)
    w = x if theta is None else warp_fwd(x, theta, in_range, dt)[0]
    if in_range == "m11":
        return normalize_minus11(alpha_composite(normalize_zero1(w)))
    return alpha_composite(w)


def render_fwd_bwd(x, theta, grad_out, in_range: str = "m11", dtype=np.float32):
    """Forward and analytic backward of ``render_fwd``.

    Returns dict(out, grad_x, grad_theta (None if theta is None), nan_mask[B,H,W]).
    """
    dt = _t(dtype)
    x = np.asarray(x, dt)
    B, L, C, H, W = x.shape
    g = np.asarray(grad_out, dt)
    if theta is None:
        w, aux = x, None
    else:
        w, aux = warp_fwd(x, theta, in_range, dt)
    if in_range == "m11":
        z = normalize_zero1(w)
        out = normalize_minus11(alpha_composite(z))
        gz, nan_mask = alpha_composite_bwd(z, (g * dt(2)).astype(dt))
        gw = (gz * dt(0.5)).astype(dt)
    else:
        z = w
        out = alpha_composite(z)
        gw, nan_mask = alpha_composite_bwd(z, g)
    if theta is None:
        return dict(out=out, grad_x=gw, grad_theta=None, nan_mask=nan_mask)
    gimg, ggx, ggy = grid_sample_bwd((B * L, C, H, W), aux, gw.reshape(B * L, C, H, W))
    gtheta = affine_grid_bwd(ggx, ggy, dt).reshape(B, L, 2, 3)
    return dict(out=out, grad_x=gimg.reshape(x.shape), grad_theta=gtheta, nan_mask=nan_mask)


# ---- the non-differentiable Pillow composite (custom_utils/image_utils.py:74-96) -----------------------------
# The integer "over" lives in a third-party dependency that is not under /root/reference: Pillow
# (libImaging/AlphaComposite.c; 12.2.0 installed in this image, the reference pins none).  Restated here from its
# published algorithm and pinned against Pillow itself, through the reference's own function, by
# tests/golden/pil_composite_golden.npz (oracle/make_golden_pil.py).

def pil_to_byte(v01) -> np.ndarray:
    """``ToPILImage`` on a float tensor: ``trunc(v * 255)`` with the product in fp32 (torchvision
    ``functional.to_pil_image``: ``pic.mul(255).byte()`` / ``(npimg * 255).astype(np.uint8)``).  Values outside
    [0,1] saturate here (the cast wraps there; the reference never feeds such values)."""
    s = np.asarray(v01, np.float32) * np.float32(255)
    return np.clip(np.trunc(np.nan_to_num(s, nan=0.0)), 0, 255).astype(np.uint8)


def pil_over(dst: np.ndarray, src: np.ndarray) -> np.ndarray:
    """``dst.alpha_composite(src)`` on uint8 RGBA arrays [...,4] (channel last), Pillow's integer arithmetic:
    7 extra precision bits, divisions by 255 as ``((a >> 8) + a) >> 8``, transparent source copies dst."""
    d = dst.astype(np.uint32)
    s = src.astype(np.uint32)
    da, sa = d[..., 3], s[..., 3]
    blend = da * (255 - sa)
    outa255 = sa * 255 + blend
    coef1 = sa * 255 * 255 * 128 // np.where(outa255 == 0, 1, outa255)
    coef2 = 255 * 128 - coef1
    out = np.empty_like(d)
    for c in range(3):
        t = s[..., c] * coef1 + d[..., c] * coef2 + (0x80 << 7)
        out[..., c] = (((t >> 8) + t) >> 8) >> 7
    t = outa255 + 0x80
    out[..., 3] = ((t >> 8) + t) >> 8
    return np.where((sa == 0)[..., None], d, out).astype(np.uint8)


def pil_alpha_composite(x, in_range: str = "01"):
    """``image_utils.alpha_composite`` on ``[B,L,4,H,W]`` (or ``[L,4,H,W]``): returns (fp32 ``byte / 255``
    ``[B,4,H,W]``, the uint8 canvas).  ``in_range='m11'`` applies ``normalize_zero1`` first (``(t + 1) / 2`` in
    fp32, ``image_utils.py:184-187``), as the callers at ``custom/loss_aio.py:351,362`` do."""
    x = np.asarray(x, np.float32)
    unb = x.ndim == 4
    if unb:
        x = x[None]
    if in_range == "m11":
        x = ((x + np.float32(1)) / np.float32(2)).astype(np.float32)
    b = np.moveaxis(pil_to_byte(x), 2, -1)                   # [B,L,H,W,4]
    canvas = b[:, 0]
    for l in range(1, b.shape[1]):
        canvas = pil_over(canvas, b[:, l])
    u8 = np.moveaxis(canvas, -1, 1)                          # [B,4,H,W]
    out = (u8.astype(np.float32) / np.float32(255)).astype(np.float32)
    return (out[0], u8[0]) if unb else (out, u8)
