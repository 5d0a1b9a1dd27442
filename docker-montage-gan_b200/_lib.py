"""ctypes binding of libmontage_render.so (the C ABI in include/montage_render.h).

Mirrors the role of the reference's plugin loader (``torch_utils/custom_ops.py:49-129`` +
``torch_utils/ops/bias_act.py:52-66``) with two deliberate differences: the library is built
ahead of time (``build.py``), and there is NO fallback -- if the library is missing or a call
fails, the caller gets an exception (the reference silently falls back to a slow reference
implementation, ``bias_act.py:63-66``; a renderer that silently ran on the CPU would void every
parity and performance claim).
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmontage_render.so")

MGR_F32, MGR_BF16, MGR_F16 = 0, 1, 2
MGR_RANGE_M11, MGR_RANGE_01 = 0, 1
MGR_NEED_GRAD_X, MGR_NEED_GRAD_THETA = 1, 2
ABI_VERSION = 4

_c = ctypes
_vp, _i, _sz = _c.c_void_p, _c.c_int, _c.c_size_t
_i64p = _c.POINTER(_c.c_int64)

class MgrLayer(_c.Structure):
    """``MgrLayer`` of include/montage_render.h: one layer of a ragged stack (or where its gradient goes)."""
    _fields_ = [("ptr", _vp), ("sb", _c.c_int64), ("sc", _c.c_int64), ("sh", _c.c_int64),
                ("h", _i), ("w", _i), ("top", _i), ("left", _i)]


_layp = _c.POINTER(MgrLayer)

# name -> (restype, argtypes); every symbol include/montage_render.h declares
SYMBOLS = {
    "mgr_abi_version": (_i, []),
    "mgr_build_info": (_c.c_char_p, []),
    "mgr_last_error": (_c.c_char_p, []),
    "mgr_kernel_launch_count": (_c.c_longlong, []),
    "mgr_set_debug_path": (_i, [_i]),
    "mgr_saved_alpha_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mgr_render_forward": (_i, [_vp, _i64p, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_render_backward_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "mgr_render_backward_workspace_bytes_for": (_sz, [_vp, _i64p, _i, _i, _i, _i, _i, _i, _i, _i]),
    "mgr_warp_forward": (_i, [_vp, _i64p, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_warp_backward_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "mgr_warp_backward": (_i, [_vp, _i64p, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_translation_to_theta": (_i, [_vp, _vp, _c.c_longlong, _vp]),
    "mgr_pad_stack_layer": (_i, [_vp, _i64p, _vp, _i, _i, _i, _i, _i, _i, _i, _c.c_float, _i, _vp]),
    "mgr_composite_jvp": (_i, [_vp, _i64p, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_render_forward_ragged": (_i, [_layp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_render_backward_ragged": (_i, [_layp, _vp, _vp, _vp, _vp, _layp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_augment_geom_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "mgr_augment_geom_forward": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_augment_geom_backward": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_composite_u8": (_i, [_vp, _i64p, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_render_host_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mgr_render_fwd_bwd_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mgr_render_backward": (_i, [_vp, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None
launch_count = 0   # number of C-ABI compute calls issued by this process (bench.py reports it)


class MontageRenderError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Load (building first if stale and nvcc is available) and type the library.  Raises
    ``MontageRenderError`` when it cannot -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            from . import build as _build
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                if not os.path.isfile(LIB_PATH):
                    raise MontageRenderError(f"libmontage_render.so is missing and could not be built: {exc}") from exc
                # a library exists but its sources changed and the rebuild failed: loading it would silently run old
                # kernels under new host code.  Opt in explicitly (MGR_ALLOW_STALE_LIB=1) to do that.
                if os.environ.get("MGR_ALLOW_STALE_LIB") != "1":
                    raise MontageRenderError("libmontage_render.so is stale (source hash mismatch) and the rebuild failed: "
                                             f"{exc}; set MGR_ALLOW_STALE_LIB=1 to load it anyway") from exc
                import warnings
                warnings.warn(f"loading a STALE libmontage_render.so (rebuild failed: {exc})", RuntimeWarning)
        if not os.path.isfile(LIB_PATH):
            raise MontageRenderError(f"{LIB_PATH} not found; run `python __graft_entry__.py build`")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise MontageRenderError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = res
            fn.argtypes = args
        got = lib.mgr_abi_version()
        if got != ABI_VERSION:
            raise MontageRenderError(f"ABI mismatch: library {got}, binding {ABI_VERSION}")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mgr_last_error().decode(errors="replace")
        raise MontageRenderError(f"{what} failed (code {rc}): {msg}")


def strides_arg(strides):
    if strides is None:
        return None
    return (_c.c_int64 * 5)(*[int(s) for s in strides])
