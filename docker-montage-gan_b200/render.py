"""Host side of the analytic renderer: ``torch.autograd.Function`` wrappers over the C ABI.

Public functions mirror the reference's own names and tensor conventions (paths relative to
/root/reference/montage_gan):

* ``render(x, theta)``                -- the chain ``STNv2c`` warp (``fukuwarai/networks.py:250-257``)
  -> ``normalize_minus11(alpha_composite_pytorch(normalize_zero1(.)))`` (``custom/loss_aio.py:251``)
  fused into one kernel; ``theta=None`` is the composite-only real branch (``loss_aio.py:313-320``).
* ``alpha_composite_pytorch(blchw)``  -- drop-in for ``custom_utils/image_utils.py:112-163``
  (default branch; [0,1] range; accepts [B,L,4,H,W] or [L,4,H,W]).

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic happens in
libmontage_render.so.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.MGR_F32, torch.bfloat16: _lib.MGR_BF16, torch.float16: _lib.MGR_F16}
_RANGES = {"m11": _lib.MGR_RANGE_M11, "01": _lib.MGR_RANGE_01}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_inputs(x, theta, in_range):
    if not isinstance(x, torch.Tensor):
        raise TypeError("x must be a torch.Tensor")
    if x.dim() != 5 or x.shape[2] != 4:
        raise ValueError(f"x must be [B,L,4,H,W] (RGBA layers), got {tuple(x.shape)}")
    if x.dtype not in _DTYPES:
        raise TypeError(f"x dtype {x.dtype} not supported (float32, bfloat16, float16)")
    if in_range not in _RANGES:
        raise ValueError(f"in_range must be 'm11' or '01', got {in_range!r}")
    B, L = x.shape[:2]
    if L < 1:
        raise ValueError("x needs at least one layer")
    if theta is not None:
        if tuple(theta.shape) != (B, L, 2, 3):
            raise ValueError(f"theta must be [B,L,2,3] = {(B, L, 2, 3)}, got {tuple(theta.shape)}")
        if theta.device != x.device:
            raise ValueError("theta must live on the same device as x")
    if not x.is_cuda:
        raise _lib.MontageRenderError("x must be a CUDA tensor: the renderer has no CPU path")


def _x_arg(x):
    """Pass x through untouched whenever the innermost stride is 1 (any batch / layer / channel /
    row strides are handled by the kernels); only a W-strided view is compacted."""
    if x.stride(4) != 1 and x.shape[4] > 1:
        x = x.contiguous()
    strides = None if x.is_contiguous() else _lib.strides_arg(x.stride()[:4] + (1,))
    return x, strides


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, in_range):
        lib = _lib.load()
        B, L, _, H, W = x.shape
        xk, strides = _x_arg(x.detach())
        th = None if theta is None else theta.detach().to(torch.float32).contiguous()
        out = torch.empty((B, 4, H, W), dtype=x.dtype, device=x.device)
        # alpha samples for the atomics-free backward: only when a gradient will be asked for
        sav = None
        if th is not None and any(ctx.needs_input_grad[:2]):
            sav = torch.empty(lib.mgr_saved_alpha_bytes(B, L, H, W, _DTYPES[x.dtype]), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mgr_render_forward(_ptr(xk), strides, _ptr(th), _ptr(out), _ptr(sav), B, L, H, W, _DTYPES[x.dtype],
                                        _RANGES[in_range], _stream_ptr(x.device))
        _lib.check(rc, "mgr_render_forward")
        _lib.launch_count += 1
        ctx.in_range = in_range
        ctx.theta_dtype = None if theta is None else theta.dtype
        ctx.save_for_backward(xk, th, out, sav)
        ctx.x_strides = strides
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xk, th, out, sav = ctx.saved_tensors
        need_x = ctx.needs_input_grad[0]
        need_t = th is not None and ctx.needs_input_grad[1]
        if not (need_x or need_t):
            return None, None, None
        if th is None:
            # composite only (the real-image branch): a Function of its own so that it can be differentiated
            # again w.r.t. grad_out -- the R1 penalty needs that (custom/loss_aio.py:327-338)
            return _CompositeBackward.apply(grad_out, xk, out, ctx.in_range, ctx.x_strides), None, None
        if torch.is_grad_enabled() and grad_out.requires_grad:
            raise NotImplementedError("double backward through the warp is not implemented (only the composite-only "
                                      "path, theta=None, is twice differentiable)")
        gx, gt = _render_backward(xk, th, out, sav, grad_out, ctx.in_range, ctx.x_strides, need_x, need_t)
        if gt is not None and ctx.theta_dtype != torch.float32:
            gt = gt.to(ctx.theta_dtype)
        return gx, gt, None


def _render_backward(xk, th, out, sav, grad_out, in_range, x_strides, need_x, need_t):
    lib = _lib.load()
    B, L, _, H, W = xk.shape
    flags = (_lib.MGR_NEED_GRAD_X if need_x else 0) | (_lib.MGR_NEED_GRAD_THETA if need_t else 0)
    go = grad_out.detach().to(xk.dtype).contiguous()
    gx = torch.empty((B, L, 4, H, W), dtype=xk.dtype, device=xk.device) if need_x else None
    gt = torch.empty((B, L, 2, 3), dtype=torch.float32, device=xk.device) if need_t else None
    dt = _DTYPES[xk.dtype]
    # what THIS tensor's path needs (the tiled kernels take about half of the upper bound for 16-bit tensors)
    ws_bytes = lib.mgr_render_backward_workspace_bytes_for(_ptr(xk), x_strides, int(th is not None), int(sav is not None),
                                                           B, L, H, W, dt, flags)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xk.device) if ws_bytes else None
    with torch.cuda.device(xk.device):
        rc = lib.mgr_render_backward(_ptr(xk), x_strides, _ptr(th), _ptr(out), _ptr(go), _ptr(sav), _ptr(gx), _ptr(gt),
                                     _ptr(ws), ws_bytes, B, L, H, W, dt, _RANGES[in_range], flags, _stream_ptr(xk.device))
    _lib.check(rc, "mgr_render_backward")
    _lib.launch_count += 1
    return gx, gt


class _CompositeBackward(torch.autograd.Function):
    """grad_x = J(x)^T grad_out for the composite; linear in grad_out, so its own backward w.r.t. grad_out is
    the JVP J(x) v (pattern: the reference's hand-written second-order ops, torch_utils/ops/bias_act.py:198-226,
    grid_sample_gradfix.py:68-88).  The second-order term w.r.t. x is not provided (nothing in the reference's
    losses uses it: the real layers are data)."""

    @staticmethod
    def forward(ctx, grad_out, xk, out, in_range, x_strides):
        gx, _ = _render_backward(xk, None, out, None, grad_out, in_range, x_strides, True, False)
        ctx.save_for_backward(xk)
        ctx.in_range, ctx.x_strides, ctx.go_dtype = in_range, x_strides, grad_out.dtype
        return gx

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, v):
        lib = _lib.load()
        (xk,) = ctx.saved_tensors
        B, L, _, H, W = xk.shape
        vt = v.to(xk.dtype).contiguous()
        jv = torch.empty((B, 4, H, W), dtype=xk.dtype, device=xk.device)
        with torch.cuda.device(xk.device):
            rc = lib.mgr_composite_jvp(_ptr(xk), ctx.x_strides, _ptr(vt), _ptr(jv), B, L, H, W, _DTYPES[xk.dtype],
                                       _RANGES[ctx.in_range], _stream_ptr(xk.device))
        _lib.check(rc, "mgr_composite_jvp")
        return jv.to(ctx.go_dtype), None, None, None, None


def render(x: torch.Tensor, theta: torch.Tensor | None = None, *, in_range: str = "m11") -> torch.Tensor:
    """Warp every layer of ``x`` [B,L,4,H,W] by ``theta`` [B,L,2,3] (``affine_grid`` +
    bilinear zero-padded ``grid_sample``, ``align_corners=False``) and alpha-over composite back
    (layer 0) to front.  Returns [B,4,H,W] straight-alpha RGBA in the same range and dtype as ``x``.

    Differentiable w.r.t. ``x`` and ``theta``.  Where the composited alpha is exactly 0 all
    gradients are 0 (the reference yields NaN there, ``image_utils.py:128-133``).
    """
    _check_inputs(x, theta, in_range)
    return _Render.apply(x, theta, in_range)


def alpha_composite_pytorch(blchw_lchw: torch.Tensor, use_premultiplied: bool = False) -> torch.Tensor:
    """Drop-in for ``custom_utils.image_utils.alpha_composite_pytorch`` (range [0,1]).

    Only the default ``use_premultiplied=False`` behaviour exists: in the reference the batched
    premultiplied branch computes and discards its result (``image_utils.py:161-163``), so the
    default path is what every caller gets.
    """
    if use_premultiplied:
        raise NotImplementedError("use_premultiplied=True is dead code in the reference (image_utils.py:161-163)")
    if blchw_lchw.dim() == 4:
        return render(blchw_lchw.unsqueeze(0), None, in_range="01")[0]
    return render(blchw_lchw, None, in_range="01")


# --------------------------------------------------------------------------------------------------
# ragged stacks: the local generators' outputs at native size, no padded canvas
# --------------------------------------------------------------------------------------------------
def _layer_array(tensors, canvas):
    """ctypes MgrLayer[L] for [B,4,h,w] tensors centred on the canvas as pad_256 centres them
    (custom_utils/image_utils.py:216-226: pad_x1 = pad_x // 2, the odd pixel goes right / down)."""
    H, W = canvas
    arr = (_lib.MgrLayer * len(tensors))()
    for l, t in enumerate(tensors):
        h, w = t.shape[2], t.shape[3]
        arr[l] = _lib.MgrLayer(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2), h, w, (H - h) // 2, (W - w) // 2)
    return arr


class _RenderRagged(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, in_range, canvas, *layers):
        lib = _lib.load()
        H, W = canvas
        L, B = len(layers), layers[0].shape[0]
        dtype, device = layers[0].dtype, layers[0].device
        xs = [t.detach() if t.stride(3) == 1 else t.detach().contiguous() for t in layers]
        th = theta.detach().to(torch.float32).contiguous()
        out = torch.empty((B, 4, H, W), dtype=dtype, device=device)
        sav = None
        if any(ctx.needs_input_grad):
            sav = torch.empty(lib.mgr_saved_alpha_bytes(B, L, H, W, _DTYPES[dtype]), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            rc = lib.mgr_render_forward_ragged(_layer_array(xs, canvas), _ptr(th), _ptr(out), _ptr(sav), B, L, H, W,
                                               _DTYPES[dtype], _RANGES[in_range], _stream_ptr(device))
        _lib.check(rc, "mgr_render_forward_ragged")
        _lib.launch_count += 1
        ctx.in_range, ctx.canvas, ctx.theta_dtype = in_range, canvas, theta.dtype
        ctx.save_for_backward(th, out, sav, *xs)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        th, out, sav, *xs = ctx.saved_tensors
        need_t = ctx.needs_input_grad[0]
        need_x = any(ctx.needs_input_grad[3:])
        if not (need_t or need_x):
            return (None,) * (3 + len(xs))
        if torch.is_grad_enabled() and grad_out.requires_grad:
            raise NotImplementedError("double backward through the warp is not implemented")
        lib = _lib.load()
        H, W = ctx.canvas
        L, B = len(xs), xs[0].shape[0]
        dtype, device = xs[0].dtype, xs[0].device
        flags = (_lib.MGR_NEED_GRAD_X if need_x else 0) | (_lib.MGR_NEED_GRAD_THETA if need_t else 0)
        go = grad_out.detach().to(dtype).contiguous()
        gxs = [torch.empty(t.shape, dtype=dtype, device=device) for t in xs] if need_x else None
        gt = torch.empty((B, L, 2, 3), dtype=torch.float32, device=device) if need_t else None
        ws_bytes = lib.mgr_render_backward_workspace_bytes(B, L, H, W, _lib.MGR_F32, 1, flags)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            rc = lib.mgr_render_backward_ragged(_layer_array(xs, ctx.canvas), _ptr(th), _ptr(out), _ptr(go), _ptr(sav),
                                                _layer_array(gxs, ctx.canvas) if need_x else None, _ptr(gt), _ptr(ws), ws_bytes,
                                                B, L, H, W, _DTYPES[dtype], _RANGES[ctx.in_range], flags, _stream_ptr(device))
        _lib.check(rc, "mgr_render_backward_ragged")
        _lib.launch_count += 1
        if gt is not None and ctx.theta_dtype != torch.float32:
            gt = gt.to(ctx.theta_dtype)
        return (gt, None, None) + (tuple(gxs) if need_x else (None,) * L)


def render_ragged(layers, theta: torch.Tensor, *, canvas=(256, 256), in_range: str = "m11") -> torch.Tensor:
    """``render(make_batch_for_pos_estimator(layers, pad_value=-1, canvas), theta)`` without building the padded
    ``[B,L,4,H,W]`` tensor: ``layers`` is the list of the L local generators' outputs ``[B,4,h_l,w_l]`` at their
    native sizes (``training/dataset_aio.py:30-83``), each centred on the canvas the way ``pad_256`` does
    (``custom_utils/image_utils.py:216-243``).  Pixels outside a layer read as the padding value (transparent black),
    are never loaded, and get no gradient; tiles that miss a layer skip it.  Differentiable w.r.t. every layer and theta.

    Layer widths and ``(W - w) // 2`` must be multiples of 4 (the reference's sizes are multiples of 32); otherwise
    the library answers ``MGR_ERR_UNSUPPORTED`` -- build the canvas with ``make_batch_for_pos_estimator`` then.
    """
    layers = list(layers)
    if len(layers) < 2:
        raise ValueError("a ragged stack needs at least two layers")
    B, H, W = layers[0].shape[0], canvas[0], canvas[1]
    for t in layers:
        if not isinstance(t, torch.Tensor) or t.dim() != 4 or t.shape[0] != B or t.shape[1] != 4:
            raise ValueError("every layer must be a [B,4,h,w] tensor with the same B")
        if t.dtype != layers[0].dtype or t.device != layers[0].device:
            raise ValueError("layers must share dtype and device")
        if t.shape[2] > H or t.shape[3] > W:
            raise ValueError(f"layer {tuple(t.shape)} is larger than the canvas {canvas}")
    if layers[0].dtype not in _DTYPES:
        raise TypeError(f"dtype {layers[0].dtype} not supported (float32, bfloat16, float16)")
    if in_range not in _RANGES:
        raise ValueError(f"in_range must be 'm11' or '01', got {in_range!r}")
    if tuple(theta.shape) != (B, len(layers), 2, 3):
        raise ValueError(f"theta must be {(B, len(layers), 2, 3)}, got {tuple(theta.shape)}")
    if not layers[0].is_cuda or theta.device != layers[0].device:
        raise _lib.MontageRenderError("layers and theta must be CUDA tensors on one device: the renderer has no CPU path")
    return _RenderRagged.apply(theta, in_range, (int(H), int(W)), *layers)


def alpha_composite(blchw_lchw: torch.Tensor, *, in_range: str = "01", return_bytes: bool = False):
    """Drop-in for ``custom_utils.image_utils.alpha_composite`` (``image_utils.py:74-96``), the non-differentiable
    Pillow composite behind snapshots, metrics and the renderer-training targets: every layer is quantised to bytes
    (``trunc(v * 255)``), composited back to front with Pillow's integer "over", and returned as ``byte / 255``.
    Bit-exact with that path (tests/golden/pil_composite_golden.npz), without its D2H copy, per-sample / per-layer
    Python loop and H2D copy.

    ``[B,L,4,H,W]`` -> ``[B,4,H,W]`` fp32 on the input's device (``[L,4,H,W]`` -> ``[4,H,W]``), range [0,1].
    ``in_range='m11'`` folds the callers' ``normalize_zero1`` (``loss_aio.py:351``) into the kernel;
    ``return_bytes=True`` additionally returns the uint8 canvas.
    """
    x = blchw_lchw.unsqueeze(0) if blchw_lchw.dim() == 4 else blchw_lchw
    _check_inputs(x, None, in_range)
    lib = _lib.load()
    B, L, _, H, W = x.shape
    xk, strides = _x_arg(x.detach())
    out = torch.empty((B, 4, H, W), dtype=torch.float32, device=x.device)
    u8 = torch.empty((B, 4, H, W), dtype=torch.uint8, device=x.device) if return_bytes else None
    with torch.cuda.device(x.device):
        rc = lib.mgr_composite_u8(_ptr(xk), strides, _ptr(out), _ptr(u8), B, L, H, W, _DTYPES[x.dtype], _RANGES[in_range],
                                  _stream_ptr(x.device))
    _lib.check(rc, "mgr_composite_u8")
    if blchw_lchw.dim() == 4:
        out, u8 = out[0], (None if u8 is None else u8[0])
    return (out, u8) if return_bytes else out


# --------------------------------------------------------------------------------------------------
# materialised warp and the caller-side helpers (rows a4, a10, a12)
# --------------------------------------------------------------------------------------------------
class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, in_range):
        lib = _lib.load()
        B, L, _, H, W = x.shape
        xk, strides = _x_arg(x.detach())
        th = theta.detach().to(torch.float32).contiguous()
        out = torch.empty((B, L, 4, H, W), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mgr_warp_forward(_ptr(xk), strides, _ptr(th), _ptr(out), B, L, H, W, _DTYPES[x.dtype],
                                      _RANGES[in_range], _stream_ptr(x.device))
        _lib.check(rc, "mgr_warp_forward")
        ctx.in_range, ctx.theta_dtype, ctx.x_strides = in_range, theta.dtype, strides
        ctx.save_for_backward(xk, th)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _lib.load()
        xk, th = ctx.saved_tensors
        B, L, _, H, W = xk.shape
        need_x, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        flags = (_lib.MGR_NEED_GRAD_X if need_x else 0) | (_lib.MGR_NEED_GRAD_THETA if need_t else 0)
        if flags == 0:
            return None, None, None
        go = grad_out.to(xk.dtype).contiguous()
        gx = torch.empty((B, L, 4, H, W), dtype=xk.dtype, device=xk.device) if need_x else None
        gt = torch.empty((B, L, 2, 3), dtype=torch.float32, device=xk.device) if need_t else None
        dt = _DTYPES[xk.dtype]
        ws_bytes = lib.mgr_warp_backward_workspace_bytes(B, L, H, W, dt, flags)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xk.device) if ws_bytes else None
        with torch.cuda.device(xk.device):
            rc = lib.mgr_warp_backward(_ptr(xk), ctx.x_strides, _ptr(th), _ptr(go), _ptr(gx), _ptr(gt), _ptr(ws), ws_bytes,
                                       B, L, H, W, dt, _RANGES[ctx.in_range], flags, _stream_ptr(xk.device))
        _lib.check(rc, "mgr_warp_backward")
        if gt is not None and ctx.theta_dtype != torch.float32:
            gt = gt.to(ctx.theta_dtype)
        return gx, gt, None


def warp(x: torch.Tensor, theta: torch.Tensor, *, in_range: str = "m11") -> torch.Tensor:
    """Materialised warp of every layer: what ``STNv2c.forward`` returns as its first output
    (``fukuwarai/networks.py:250-257``; ``in_range='01'`` is the ``STNv2b`` / ``random_position`` form).
    ``render(x, theta) == composite(warp(x, theta))`` up to rounding."""
    if theta is None:
        raise ValueError("warp needs theta")
    _check_inputs(x, theta, in_range)
    return _Warp.apply(x, theta, in_range)


class _TranslationToTheta(torch.autograd.Function):
    @staticmethod
    def forward(ctx, translation):
        lib = _lib.load()
        tr = translation.detach().to(torch.float32).contiguous()
        theta = torch.empty(tr.shape[:-1] + (2, 3), dtype=torch.float32, device=tr.device)
        with torch.cuda.device(tr.device):
            rc = lib.mgr_translation_to_theta(_ptr(tr), _ptr(theta), tr.numel() // 2, _stream_ptr(tr.device))
        _lib.check(rc, "mgr_translation_to_theta")
        ctx.in_dtype = translation.dtype
        return theta

    @staticmethod
    def backward(ctx, grad_theta):
        return grad_theta[..., 2].to(ctx.in_dtype)


def convert_translate_to_2x3(translation_bl2_l2: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``custom_utils.image_utils.convert_translate_to_2x3`` (``:316-335``): [...,2] (dx, dy)
    -> [...,2,3], one kernel instead of a B*L Python loop of host-to-device copies."""
    if translation_bl2_l2.shape[-1] != 2:
        raise ValueError(f"translation must end in a dimension of 2, got {tuple(translation_bl2_l2.shape)}")
    if not translation_bl2_l2.is_cuda:
        raise _lib.MontageRenderError("translation must be a CUDA tensor: no CPU path")
    return _TranslationToTheta.apply(translation_bl2_l2)


def random_position(blchw: torch.Tensor, generator: torch.Generator | None = None) -> torch.Tensor:
    """Drop-in for ``custom_utils.image_utils.random_position`` (``:281-294``): every layer moved by a
    translation drawn from U(-1, 1); input in [0,1]."""
    B, L = blchw.shape[:2]
    tr = torch.empty((B, L, 2), device=blchw.device).uniform_(-1.0, 1.0, generator=generator)
    return warp(blchw, convert_translate_to_2x3(tr), in_range="01")


def make_batch_for_pos_estimator(list_of_bchw, pad_value=0, canvas=(256, 256)) -> torch.Tensor:
    """Drop-in for ``custom_utils.image_utils.make_batch_for_pos_estimator`` (``:229-243``): the L local
    generator outputs [B,4,h_l,w_l] are centre-padded to the canvas (256x256 in the reference,
    ``pad_256`` ``:216-226``) and written straight into one [B,L,4,H,W] tensor -- no per-sample F.pad
    loop, no stack/transpose/contiguous copies.  Differentiable (the backward is a crop)."""
    return _PadStack.apply(float(pad_value), tuple(canvas), *list_of_bchw)


class _PadStack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pad_value, canvas, *layers):
        lib = _lib.load()
        if not layers:
            raise ValueError("need at least one layer")
        H, W = canvas
        B, dtype, device = layers[0].shape[0], layers[0].dtype, layers[0].device
        L = len(layers)
        out = torch.empty((B, L, 4, H, W), dtype=dtype, device=device)
        ctx.boxes = []
        with torch.cuda.device(device):
            for l, t in enumerate(layers):
                if t.dim() != 4 or t.shape[0] != B or t.shape[1] != 4 or t.dtype != dtype or t.device != device:
                    raise ValueError(f"layer {l}: expected [B={B},4,h,w] {dtype} on {device}, got {tuple(t.shape)} {t.dtype}")
                h, w = t.shape[2:]
                if h > H or w > W:
                    raise ValueError(f"layer {l} ({h}x{w}) is larger than the canvas {H}x{W}")
                td = t.detach()
                strides = (ctypes.c_int64 * 4)(*td.stride())
                rc = lib.mgr_pad_stack_layer(_ptr(td), strides, _ptr(out), B, L, l, h, w, H, W, pad_value,
                                             _DTYPES[dtype], _stream_ptr(device))
                _lib.check(rc, "mgr_pad_stack_layer")
                ctx.boxes.append(((H - h) // 2, (W - w) // 2, h, w))
        return out

    @staticmethod
    def backward(ctx, grad):
        grads = [grad[:, l, :, top:top + h, left:left + w] for l, (top, left, h, w) in enumerate(ctx.boxes)]
        return (None, None, *grads)
