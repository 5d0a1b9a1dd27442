"""The translation kernels on TMA box copies (render_shift_tma.cuh / render_shift_tma_bwd.cuh) against the fp64 oracle and
against the staged stencil kernels they replace (mgr_set_debug_path(4) switches the TMA kernels off): every alignment of
the box inside a row (16-byte rule on the innermost coordinate), tiles that cross the image border in both range modes
(the [-1,1] mode patches the zero-filled part of the box), canvases that are not multiples of the tile, shifts that push
a layer off the canvas (zero-filled gradient texels), whole-pixel shifts.  Reference: fukuwarai/networks.py:246-257,
custom_utils/image_utils.py:112-163, 316-335."""
import numpy as np
import pytest
import torch

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import _lib, render as mr, synth
from oracle import restatement as R
from helpers import FWD_TOL, GRAD_TOL, max_abs, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _theta(B, L, seed, scale=1.0, px=None, H=None, W=None):
    g = torch.Generator().manual_seed(seed)
    th = torch.eye(2, 3).expand(B, L, 2, 3).clone()
    th[..., 2] = (torch.rand(B, L, 2, generator=g) * 2 - 1) * scale
    if px is not None:                              # shifts given in pixels: (dx_px, dy_px) per layer, cycled
        for l in range(L):
            th[:, l, 0, 2] = px[l % len(px)][0] * 2.0 / W
            th[:, l, 1, 2] = px[l % len(px)][1] * 2.0 / H
    return th


def _run(x, th, go, in_range, dtype, path):
    lib = _lib.load()
    _lib.check(lib.mgr_set_debug_path(path), "mgr_set_debug_path")
    try:
        xd = x.to(DEV, dtype).requires_grad_(True)
        td = th.to(DEV).requires_grad_(True)
        out = mr.render(xd, td, in_range=in_range)
        out.backward(go.to(DEV, dtype))
        torch.cuda.synchronize()
    finally:
        lib.mgr_set_debug_path(0)
    return out.detach().float().cpu().numpy(), xd.grad.float().cpu().numpy(), td.grad.cpu().numpy()


@pytest.mark.parametrize("in_range", ["m11", "01"])
@pytest.mark.parametrize("shape", [(2, 7, 64, 128), (1, 5, 40, 72), (2, 3, 100, 200), (1, 9, 16, 64)])
def test_tma_translation_vs_oracle_fp32(shape, in_range):
    """fp32 tensors: TMA forward and backward (transmittances parked in the workspace) against the fp64 oracle."""
    B, L, H, W = shape
    x = synth.make_layers(B, L, H, W, "S", seed=71)
    if in_range == "01":
        x = (x + 1) / 2
    th = _theta(B, L, 71, 0.9)
    go = synth.make_grad_out(B, H, W, seed=71)
    out, gx, gt = _run(x, th, go, in_range, torch.float32, 0)
    r64 = R.render_fwd_bwd(x.numpy(), th.numpy(), go.numpy(), in_range, np.float64)
    assert max_abs(out, r64["out"]) < FWD_TOL
    assert rel_err(gx, r64["grad_x"]) < GRAD_TOL
    old = _run(x, th, go, in_range, torch.float32, 4)        # the staged kernels on the same stack
    assert max_abs(out, old[0]) <= 1e-6 and rel_err(gx, old[1]) <= 1e-5 and rel_err(gt, old[2]) <= 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("in_range", ["m11", "01"])
def test_tma_translation_16bit_vs_oracle_and_staged(dtype, in_range):
    """16-bit tensors take the TMA kernels in both directions.  Judged against the fp64 oracle on the rounded inputs (storage
    rounding of the results + fp16 transmittances bound the error) and against the staged stencil kernels."""
    B, L, H, W = 3, 7, 80, 136                         # neither dimension a multiple of the tiles (63 x 15 anchors, 64 x 32 pixels)
    x = synth.make_layers(B, L, H, W, "S", seed=72)
    if in_range == "01":
        x = (x + 1) / 2
    x = x.to(dtype).float()
    th = _theta(B, L, 72, 0.7)
    # a back layer that covers the canvas (up to a sub-pixel band): where the composited alpha is ~0 the backward divides by it,
    # and with 16-bit `out` the theta gradient is then noise in ANY implementation (the staged kernels return the same
    # numbers to five digits on such stacks, both far from the fp64 oracle -- tools/dbg_theta.py).  Not a whole-pixel shift: grad_theta is one-sided there.
    th[:, 0, 0, 2] = 0.37 * 2 / W
    th[:, 0, 1, 2] = -0.41 * 2 / H
    go = synth.make_grad_out(B, H, W, seed=72).to(dtype).float()
    new = _run(x, th, go, in_range, dtype, 0)
    old = _run(x, th, go, in_range, dtype, 4)
    r64 = R.render_fwd_bwd(x.numpy(), th.numpy(), go.numpy(), in_range, np.float64)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11     # half an ulp of the storage type at 1
    assert max_abs(new[0], r64["out"]) < 2.5 * eps                  # outputs live in [-1, 1] / [0, 1]
    assert rel_err(new[1], r64["grad_x"]) < 3 * eps
    assert rel_err(new[2], r64["grad_theta"]) < 2e-2
    # the two implementations round at the same places except the transmittance (fp16 in shared memory here)
    assert max_abs(new[0], old[0]) <= 2 * eps
    assert rel_err(new[1], old[1]) <= 3 * eps
    assert rel_err(new[2], old[2]) <= 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tma_every_alignment_and_border_case(dtype):
    """Whole-pixel and fractional shifts chosen so that the box origin takes every residue modulo 8 texels, to both sides of
    the canvas, including layers pushed off it entirely: results must match the staged kernels (same arithmetic: the
    forward bit for bit) and untouched gradient texels must be exactly zero."""
    B, H, W = 2, 48, 128
    px = [(0, 0), (1, 0), (-1, 2), (2.5, -3), (3, 1), (-4.25, 0), (5, -7), (6.5, 9), (-7, 11), (8, -15), (-9.75, 16),
          (37, -5), (-63, 20), (64, -47), (-128, 0), (130, 3), (0, 48), (-3, -49), (127.5, 47.5), (-127.5, -47.5)]
    L = len(px)
    x = synth.make_layers(B, L, H, W, "S", seed=73).to(dtype).float()
    th = _theta(B, L, 73, px=px, H=H, W=W)
    go = synth.make_grad_out(B, H, W, seed=73).to(dtype).float()
    new = _run(x, th, go, "m11", dtype, 0)
    old = _run(x, th, go, "m11", dtype, 4)
    if dtype == torch.float32:
        assert max_abs(new[0], old[0]) <= 1e-6
        assert rel_err(new[1], old[1]) <= 1e-6
    else:
        assert max_abs(new[0], old[0]) <= 2.0 ** -7
        assert rel_err(new[1], old[1]) <= 2.0 ** -6
    assert rel_err(new[2], old[2]) <= (1e-4 if dtype == torch.float32 else 2e-2)
    # texels no pixel samples: exactly zero gradient, written (not left over from the allocation)
    for l in range(L):                                          # the integer part of the shift as the kernels see it (fp32 theta)
        X = int(np.floor(float(th[0, l, 0, 2].double() * 0.5 * W)))
        Y = int(np.floor(float(th[0, l, 1, 2].double() * 0.5 * H)))
        cols = np.arange(W)
        rows = np.arange(H)
        dead_c = (cols < X) | (cols > X + W)
        dead_r = (rows < Y) | (rows > Y + H)
        g = new[1][:, l]
        assert not g[..., dead_c].any(), (l, "columns")
        assert not g[:, :, dead_r, :].any(), (l, "rows")


def test_tma_16bit_many_layers_uses_the_workspace_for_transmittances():
    """More than 19 layers of 16-bit tensors: the transmittances no longer fit shared memory next to two CTAs per SM and are
    parked in the workspace instead (the fp32 policy, kGlobalT); up to 32 layers stay on the box-copy kernels."""
    for L in (21, 32):
        B, H, W = 2, 40, 72
        x = synth.make_layers(B, L, H, W, "S", seed=76).to(torch.bfloat16).float()
        th = _theta(B, L, 76, 0.6)
        th[:, 0, 0, 2] = 0.37 * 2 / W
        th[:, 0, 1, 2] = -0.41 * 2 / H
        go = synth.make_grad_out(B, H, W, seed=76).to(torch.bfloat16).float()
        new = _run(x, th, go, "m11", torch.bfloat16, 0)
        old = _run(x, th, go, "m11", torch.bfloat16, 4)
        r64 = R.render_fwd_bwd(x.numpy(), th.numpy(), go.numpy(), "m11", np.float64)
        assert max_abs(new[0], r64["out"]) < 2.5 * 2.0 ** -8
        assert rel_err(new[1], r64["grad_x"]) < 3 * 2.0 ** -8
        assert rel_err(new[1], old[1]) <= 3 * 2.0 ** -8 and rel_err(new[2], old[2]) <= 2e-2


def test_tma_backward_is_deterministic():
    B, L, H, W = 4, 7, 64, 192
    x = synth.make_layers(B, L, H, W, "S", seed=74)
    th = _theta(B, L, 74, 1.0)
    go = synth.make_grad_out(B, H, W, seed=74)
    a = _run(x, th, go, "m11", torch.bfloat16, 0)
    b = _run(x, th, go, "m11", torch.bfloat16, 0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_tma_mixed_batch_and_fallbacks():
    """A batch that mixes translation samples with general placements (the kernels partition it on the per-sample flag),
    and a row stride that breaks TMA's 16-byte rule (a sliced view: the staged stencil kernel takes over silently)."""
    B, L, H, W = 4, 5, 64, 64
    x = synth.make_layers(B, L, H, W, "S", seed=75).to(torch.bfloat16).float()
    th = _theta(B, L, 75, 0.5)
    th[2] = synth.make_theta(1, L, "I", seed=75)[0]
    go = synth.make_grad_out(B, H, W, seed=75).to(torch.bfloat16).float()
    new = _run(x, th, go, "m11", torch.bfloat16, 0)
    old = _run(x, th, go, "m11", torch.bfloat16, 4)
    assert np.array_equal(new[0][2], old[0][2]) and np.array_equal(new[1][2], old[1][2])     # the general sample: same kernels
    assert max_abs(new[0], old[0]) <= 2.0 ** -7 and rel_err(new[1], old[1]) <= 2.0 ** -6
    # W + 4 columns allocated, the first W used: rows start every 2 * (W + 4) bytes -- not a multiple of 16
    xw = torch.full((B, L, 4, H, W + 4), -1.0).to(DEV, torch.bfloat16)
    xw[..., :W] = x.to(DEV, torch.bfloat16)
    xv = xw[..., :W].requires_grad_(True)
    out = mr.render(xv, th.to(DEV), in_range="m11")
    out.backward(go.to(DEV, torch.bfloat16))
    assert max_abs(out.detach().float().cpu().numpy(), old[0]) == 0.0
    # W + 8 columns allocated: rows start every 2 * (W + 8) bytes, a multiple of 16 -- the tensor map takes the strides as they
    # are and the view stays on the box-copy kernels: same bits as the contiguous tensor
    xw8 = torch.full((B, L, 4, H + 3, W + 8), -1.0).to(DEV, torch.bfloat16)
    xw8[..., :H, :W] = x.to(DEV, torch.bfloat16)
    xv8 = xw8[..., :H, :W].requires_grad_(True)
    out8 = mr.render(xv8, th.to(DEV), in_range="m11")
    out8.backward(go.to(DEV, torch.bfloat16))
    assert max_abs(out8.detach().float().cpu().numpy(), new[0]) == 0.0
    assert max_abs(xv8.grad.float().cpu().numpy(), new[1]) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("in_range", ["m11", "01"])
def test_tma_materialised_warp_of_translation_layers(dtype, in_range):
    """mgr_warp_forward (what STNv2c returns, fukuwarai/networks.py:250-257): layers that are pure translations take the
    box-copy kernel, the others the staged one, in the same call; every box alignment, layers pushed off the canvas."""
    B, H, W = 2, 80, 136
    px = [(0, 0), (1.5, -2.25), (-3.25, 7.5), (5, 5), (-6.5, -9.75), (7, 0), (140, 3), (-20, -90), (63.5, 31.5)]
    frac = [1, 2, 4, 8]                                         # layers whose shift is fractional along both axes
    L = len(px) + 1
    x = synth.make_layers(B, L, H, W, "S", seed=77)
    if in_range == "01":
        x = (x + 1) / 2
    x = x.to(dtype).float()
    th = _theta(B, L, 77, px=px, H=H, W=W)
    th[:, L - 1] = synth.make_theta(B, 1, "I", seed=77, cover_back=False)[:, 0]     # one general layer among the translations
    gw = torch.randn(B, L, 4, H, W, generator=torch.Generator().manual_seed(77)).to(dtype).float()
    lib = _lib.load()
    res, grads = [], []
    for path in (0, 4):
        _lib.check(lib.mgr_set_debug_path(path), "mgr_set_debug_path")
        try:
            xd = x.to(DEV, dtype).requires_grad_(True)
            td = th.to(DEV).requires_grad_(True)
            w = mr.warp(xd, td, in_range=in_range)
            w.backward(gw.to(DEV, dtype))
            res.append(w.detach().float().cpu().numpy())
            grads.append((xd.grad.float().cpu().numpy(), td.grad.cpu().numpy()))
        finally:
            lib.mgr_set_debug_path(0)
    ref, aux = R.warp_fwd(x.numpy(), th.numpy(), in_range, np.float64)
    gimg, _, _ = R.grid_sample_bwd((B * L, 4, H, W), aux, gw.numpy().astype(np.float64).reshape(B * L, 4, H, W))
    f32 = dtype == torch.float32
    tol = 5e-6 if f32 else 2.0 ** -7                          # (the staged kernel carries per-pixel fp32 coordinates, the stencil one layer-wide weights)
    assert max_abs(res[0], res[1]) <= tol
    assert np.array_equal(res[0][:, L - 1], res[1][:, L - 1])                       # the general layer: the same kernel either way
    assert max_abs(res[0], ref) <= (FWD_TOL if f32 else 2.0 ** -7)
    # backward: grad_x of the translation layers is the same kernel run as its own adjoint on the upstream gradient
    assert rel_err(grads[0][0], gimg.reshape(x.shape)) <= (GRAD_TOL if f32 else 2.0 ** -7)
    assert rel_err(grads[0][0], grads[1][0]) <= (1e-5 if f32 else 2.0 ** -7)
    assert np.array_equal(grads[0][0][:, L - 1], grads[1][0][:, L - 1])
    # grad_theta (box-copy kernel for the translation layers): against the oracle and the staged kernel where the derivative is
    # two-sided (at a whole-pixel shift the cell the one-sided derivative is taken in depends on the last bit of ix)
    _, ggx, ggy = R.grid_sample_bwd((B * L, 4, H, W), aux, gw.numpy().astype(np.float64).reshape(B * L, 4, H, W))
    gth = R.affine_grid_bwd(ggx, ggy, np.float64).reshape(B, L, 2, 3)
    assert rel_err(grads[0][1][:, frac], gth[:, frac]) <= (5e-4 if f32 else 2e-2)
    assert rel_err(grads[0][1][:, frac], grads[1][1][:, frac]) <= (5e-4 if f32 else 2e-2)
    assert rel_err(grads[0][1][:, L - 1], grads[1][1][:, L - 1]) <= 1e-5            # the general layer: the same kernel, atomics order only
