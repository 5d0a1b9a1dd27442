"""Ragged stacks: one [B,4,h,w] tensor per layer at its native size, centred on the canvas, instead
of the padded [B,L,4,H,W] tensor of make_batch_for_pos_estimator (custom_utils/image_utils.py:216-243).

The semantics are those of padding with the transparent value and rendering the canvas, so the checks are
(1) parity with the oracle run on the padded canvas (three-way rule of tests/helpers.py), (2) agreement with this
library's own canvas path on the padded tensor -- forward bit for bit: the same texel values reach the same
arithmetic --, and (3) the argument rules of the C ABI."""
import numpy as np
import pytest
import torch

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import _lib, render as mr, synth
from oracle import restatement as R
from helpers import FWD_TOL, GRAD_TOL, max_abs, rel_err, three_way

pytestmark = pytest.mark.gpu

# the reference's nine face-part sizes (custom/dataset_aio.py:28-83), h x w on a 256 x 256 canvas
REFERENCE_SIZES = [(256, 256), (256, 256), (160, 224), (256, 256), (96, 160), (64, 96), (64, 32), (256, 256), (64, 160)]


def _make(B, canvas, sizes, family, seed, dtype=torch.float32, pad=-1.0):
    H, W = canvas
    full = synth.make_layers(B, len(sizes), H, W, family, seed=seed)
    layers, padded = [], torch.full((B, len(sizes), 4, H, W), pad)
    for l, (h, w) in enumerate(sizes):
        top, left = (H - h) // 2, (W - w) // 2
        t = full[:, l, :, top:top + h, left:left + w].contiguous().to(dtype)
        layers.append(t)
        padded[:, l, :, top:top + h, left:left + w] = t.float()
    return layers, padded.to(dtype)


def _crop(g, sizes, canvas):
    H, W = canvas
    return [g[:, l, :, (H - h) // 2:(H - h) // 2 + h, (W - w) // 2:(W - w) // 2 + w] for l, (h, w) in enumerate(sizes)]


def _run_ragged(layers, theta, go, canvas, in_range="m11"):
    xs = [t.cuda().requires_grad_(True) for t in layers]
    th = theta.cuda().requires_grad_(True)
    out = mr.render_ragged(xs, th, canvas=canvas, in_range=in_range)
    grads = torch.autograd.grad(out, xs + [th], go.cuda().to(out.dtype))
    return out.detach().cpu(), [g.cpu() for g in grads[:-1]], grads[-1].cpu()


def _run_canvas(padded, theta, go, in_range="m11"):
    x = padded.cuda().requires_grad_(True)
    th = theta.cuda().requires_grad_(True)
    out = mr.render(x, th, in_range=in_range)
    gx, gt = torch.autograd.grad(out, (x, th), go.cuda().to(out.dtype))
    return out.detach().cpu(), gx.cpu(), gt.cpu()


CASES = [
    ((64, 64), [(64, 64), (40, 56), (24, 40), (17, 24), (16, 8)]),
    ((96, 80), [(96, 80), (33, 48), (96, 16), (8, 80)]),
    ((64, 128), [(20, 64), (64, 128), (64, 8)]),
]


@pytest.mark.parametrize("canvas,sizes", CASES)
@pytest.mark.parametrize("tf", ["I", "T", "X"])
def test_ragged_parity_vs_oracle_on_padded_canvas(canvas, sizes, tf):
    B, L = 2, len(sizes)
    layers, padded = _make(B, canvas, sizes, "S", seed=21)
    theta = synth.make_theta(B, L, tf, seed=21, cover_back=(sizes[0] == canvas))
    go = synth.make_grad_out(B, canvas[0], canvas[1], "randn", seed=21)
    out, gxs, gt = _run_ragged(layers, theta, go, canvas)
    r32 = R.render_fwd_bwd(padded.numpy(), theta.numpy(), go.numpy(), "m11", np.float32)
    r64 = R.render_fwd_bwd(padded.numpy(), theta.numpy(), go.numpy(), "m11", np.float64)
    ok, info = three_way(out.numpy(), r32["out"], r64["out"], FWD_TOL, max_abs)
    assert ok, info
    got = torch.zeros(padded.shape)
    inside = torch.zeros(padded.shape, dtype=torch.bool)
    for l, (g, (h, w)) in enumerate(zip(gxs, sizes)):
        top, left = (canvas[0] - h) // 2, (canvas[1] - w) // 2
        got[:, l, :, top:top + h, left:left + w] = g
        inside[:, l, :, top:top + h, left:left + w] = True
    m = inside.numpy()                                                 # the padding has no gradient in the ragged form
    ok, info = three_way(np.where(m, got.numpy(), 0), np.where(m, r32["grad_x"], 0), np.where(m, r64["grad_x"], 0), GRAD_TOL, rel_err)
    assert ok, info
    if tf != "X":                                                      # grad_theta is ill-conditioned for extreme placements
        ok, info = three_way(gt.numpy(), r32["grad_theta"], r64["grad_theta"], 2e-3, rel_err)
        assert ok, info


@pytest.mark.parametrize("canvas,sizes", CASES + [((256, 256), REFERENCE_SIZES)])
@pytest.mark.parametrize("tf", ["I", "T"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_ragged_equals_canvas_path_on_padded_tensor(canvas, sizes, tf, dtype):
    B, L = 2, len(sizes)
    layers, padded = _make(B, canvas, sizes, "F" if dtype == torch.float32 else "S", seed=5, dtype=dtype)
    theta = synth.make_theta(B, L, tf, seed=5, cover_back=False)
    go = synth.make_grad_out(B, canvas[0], canvas[1], "randn", seed=5)
    out, gxs, gt = _run_ragged(layers, theta, go, canvas)
    out_c, gx_c, gt_c = _run_canvas(padded, theta, go)
    assert torch.equal(out, out_c)                                      # same texels, same arithmetic
    # canvases of pure translations take the backward on TMA box copies (render_shift_tma_bwd.cuh: another summation order,
    # and fp16 transmittances for 16-bit tensors), ragged stacks the staged stencil kernel: equal to rounding there, bit for
    # bit everywhere else
    tma = tf == "T"
    # (fp32: these stacks have no covering layer, so G_P = g / A is formed with A down to ~1e-3 and two summation orders part by 1e-4)
    tol = {torch.float32: 5e-4, torch.bfloat16: 2.0 ** -6, torch.float16: 2.0 ** -9}[dtype]
    for g, gc in zip(gxs, _crop(gx_c, sizes, canvas)):
        if tma:
            assert rel_err(g.float().numpy(), gc.float().numpy()) <= tol
        else:
            assert torch.equal(g, gc)
    assert rel_err(gt.numpy(), gt_c.numpy()) < ((2e-3 if dtype == torch.float32 else 2e-2) if tma else 1e-5)  # atomics: summation order only


def test_ragged_range01_and_partial_grads():
    canvas, sizes = (64, 64), [(64, 64), (32, 32), (16, 48)]
    layers, padded = _make(2, canvas, sizes, "W", seed=9, pad=0.0)
    layers = [(t + 1) / 2 for t in layers]
    padded = torch.where(padded == 0, padded, (padded + 1) / 2)
    for l, (h, w) in enumerate(sizes):                                   # rebuild the padded canvas exactly
        padded[:, l] = 0
        padded[:, l, :, (64 - h) // 2:(64 - h) // 2 + h, (64 - w) // 2:(64 - w) // 2 + w] = layers[l]
    theta = synth.make_theta(2, 3, "I", seed=9)
    go = synth.make_grad_out(2, 64, 64, "randn", seed=9)
    out, gxs, gt = _run_ragged(layers, theta, go, canvas, in_range="01")
    out_c, gx_c, gt_c = _run_canvas(padded, theta, go, in_range="01")
    assert torch.equal(out, out_c)
    for g, gc in zip(gxs, _crop(gx_c, sizes, canvas)):
        assert torch.equal(g, gc)
    # theta-only and layers-only gradients
    xs = [t.cuda() for t in layers]
    th = theta.cuda().requires_grad_(True)
    (g_only_t,) = torch.autograd.grad(mr.render_ragged(xs, th, canvas=canvas, in_range="01"), th, go.cuda())
    assert rel_err(g_only_t.cpu().numpy(), gt.numpy()) < 1e-5
    xs = [t.cuda().requires_grad_(True) for t in layers]
    g_only_x = torch.autograd.grad(mr.render_ragged(xs, theta.cuda(), canvas=canvas, in_range="01"), xs, go.cuda())
    for a, b in zip(g_only_x, gxs):
        assert torch.equal(a.cpu(), b)


def test_ragged_argument_rules():
    canvas = (64, 64)
    good = [torch.zeros(2, 4, 64, 64, device="cuda"), torch.zeros(2, 4, 16, 16, device="cuda")]
    theta = synth.make_theta(2, 2, "T", seed=0).cuda()
    mr.render_ragged(good, theta, canvas=canvas)
    with pytest.raises(_lib.MontageRenderError, match="multiples of 4"):
        mr.render_ragged([good[0], torch.zeros(2, 4, 16, 18, device="cuda")], theta, canvas=canvas)     # width 18
    with pytest.raises(_lib.MontageRenderError, match="multiples of 4"):
        mr.render_ragged([good[0], torch.zeros(2, 4, 16, 60, device="cuda")], theta, canvas=canvas)     # left = 2
    with pytest.raises(ValueError):
        mr.render_ragged([good[0]], theta[:, :1], canvas=canvas)                                         # one layer
    with pytest.raises(ValueError):
        mr.render_ragged([good[0], torch.zeros(2, 4, 80, 16, device="cuda")], theta, canvas=canvas)     # taller than the canvas
    with pytest.raises(_lib.MontageRenderError):
        mr.render_ragged([t.cpu() for t in good], theta.cpu(), canvas=canvas)                            # no CPU path


def test_ragged_layers_may_be_views_into_a_padded_canvas():
    """A caller that already holds the padded [B,L,4,H,W] tensor (the STN's localisation CNN reads it) can still tell
    the renderer where each layer's content is: strided views of the canvas are valid ragged layers, nothing is copied."""
    canvas, sizes = (128, 128), [(128, 128), (64, 96), (32, 48), (17, 64)]
    B, L = 2, len(sizes)
    _, padded = _make(B, canvas, sizes, "S", seed=13)
    theta = synth.make_theta(B, L, "I", seed=13)
    go = synth.make_grad_out(B, 128, 128, "randn", seed=13)
    xp = padded.cuda()
    views = [xp[:, l, :, (128 - h) // 2:(128 - h) // 2 + h, (128 - w) // 2:(128 - w) // 2 + w].requires_grad_(True)
             for l, (h, w) in enumerate(sizes)]
    assert not views[1].is_contiguous()
    th = theta.cuda().requires_grad_(True)
    out = mr.render_ragged(views, th, canvas=canvas)
    grads = torch.autograd.grad(out, views + [th], go.cuda())
    out_c, gx_c, gt_c = _run_canvas(padded, theta, go)
    assert torch.equal(out.detach().cpu(), out_c)
    for g, gc in zip(grads[:-1], _crop(gx_c, sizes, canvas)):
        assert g.is_contiguous() and torch.equal(g.cpu(), gc)
    assert rel_err(grads[-1].cpu().numpy(), gt_c.numpy()) < 1e-5
