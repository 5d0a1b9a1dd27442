"""CPU tests of the drop-in boundary: libmontage_render.so loads without a GPU, exports every
symbol include/montage_render.h declares, and validates arguments with error codes (no compute)."""
import ctypes
import os
import re

import pytest

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(mgr_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no symbols parsed from include/*.h"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.mgr_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.mgr_build_info()


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "montage_render.h")).read()
    for cite in ("fukuwarai/networks.py:247-258", "custom_utils/image_utils.py:112-163", "custom/loss_aio.py:245-257"):
        assert cite in text


def test_sass_is_sm100a():
    so = _lib.LIB_PATH
    out = os.popen(f"cuobjdump -lelf {so} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out


@pytest.mark.parametrize("args,code", [
    (dict(x=None), 1),
    (dict(L=0), 1),
    (dict(H=0), 1),
    (dict(dtype=7), 1),
    (dict(range_mode=5), 1),
    (dict(L=33), 2),
    (dict(out=None), 1),
])
def test_forward_argument_validation(args, code):
    lib = _lib.load()
    a = dict(x=0x1000, theta=None, out=0x2000, B=1, L=2, H=4, W=4, dtype=0, range_mode=0)
    a.update(args)
    rc = lib.mgr_render_forward(a["x"], None, a["theta"], a["out"], None, a["B"], a["L"], a["H"], a["W"], a["dtype"],
                                a["range_mode"], None)
    assert rc == code
    assert lib.mgr_last_error()


def test_strided_w_rejected():
    lib = _lib.load()
    strides = (ctypes.c_int64 * 5)(512, 128, 32, 8, 2)
    rc = lib.mgr_render_forward(0x1000, strides, None, 0x2000, None, 1, 2, 4, 4, 0, 0, None)
    assert rc == 2 and b"stride" in lib.mgr_last_error()


def test_backward_validation_and_workspace():
    lib = _lib.load()
    assert lib.mgr_render_backward_workspace_bytes(2, 3, 8, 8, _lib.MGR_F32, 0, 3) == 0      # composite only
    assert lib.mgr_render_backward_workspace_bytes(2, 3, 8, 8, _lib.MGR_BF16, 0, 1) == 0
    assert lib.mgr_render_backward_workspace_bytes(2, 3, 8, 8, _lib.MGR_F32, 1, 3) == 2 * 3 * 64 * 8 + 2 * 64 * 16 + 2 * 3 * (128 + 4 + 4) + 32 + 2 * 4   # records, G_P, plans + order + work list, counters, sample flags
    assert lib.mgr_saved_alpha_bytes(2, 3, 8, 8, _lib.MGR_F32) == 2 * 3 * 64 * 4 + 2 * 4       # alpha samples (a multiple of 256 here) + a flag per sample
    assert lib.mgr_saved_alpha_bytes(2, 3, 8, 8, _lib.MGR_BF16) == 2 * 3 * 64 * 2 + 2 * 4
    assert lib.mgr_saved_alpha_bytes(1, 3, 8, 8, _lib.MGR_BF16) == 512 + 4                      # 384 bytes of samples, padded to 256
    need = lib.mgr_render_backward_workspace_bytes(2, 3, 8, 8, _lib.MGR_BF16, 1, 3)
    assert need == 2 * 3 * 4 * 64 * 4                # fp32 scatter accumulator of the general path dominates
    # per-tensor query: an aligned bf16 tensor with saved alphas takes the tiled kernels (records only); without saved
    # alphas, or from an odd address, the scatter path and its fp32 accumulator
    tiled = 2 * 3 * 64 * 8 + 2 * 64 * 16 + 2 * 3 * (128 + 4 + 4) + 32 + 2 * 4
    assert lib.mgr_render_backward_workspace_bytes_for(0x1000, None, 1, 1, 2, 3, 8, 8, _lib.MGR_BF16, 3) == tiled
    assert lib.mgr_render_backward_workspace_bytes_for(0x1000, None, 1, 0, 2, 3, 8, 8, _lib.MGR_BF16, 3) == need
    assert lib.mgr_render_backward_workspace_bytes_for(0x1002, None, 1, 1, 2, 3, 8, 8, _lib.MGR_BF16, 3) == need
    assert lib.mgr_render_backward_workspace_bytes_for(0x1000, None, 0, 1, 2, 3, 8, 8, _lib.MGR_BF16, 3) == 0
    rc = lib.mgr_render_backward(0x1000, None, 0x3000, 0x2000, 0x2000, None, 0x4000, 0x5000, None, 0, 2, 3, 8, 8,
                                 _lib.MGR_BF16, 0, 3, None)
    assert rc == 3 and b"workspace" in lib.mgr_last_error()
    # a warped stack of more than 65535 layers is refused by the forward already (the backward could not take it)
    rc = lib.mgr_render_forward(0x1000, None, 0x3000, 0x2000, None, 40000, 2, 4, 4, 0, 0, None)
    assert rc == 2 and b"65535" in lib.mgr_last_error()
    rc = lib.mgr_render_backward(0x1000, None, None, 0x2000, 0x2000, None, None, None, None, 0, 2, 3, 8, 8, 0, 0, 1, None)
    assert rc == 1                                   # grad_x requested but NULL
    rc = lib.mgr_render_backward(0x1000, None, None, 0x2000, 0x2000, None, None, None, None, 0, 2, 3, 8, 8, 0, 0, 0, None)
    assert rc == 0                                   # nothing requested: no-op
    rc = lib.mgr_render_forward(0x1000, None, None, 0x2000, None, 0, 3, 8, 8, 0, 0, None)
    assert rc == 0                                   # empty batch: no-op


def test_new_entry_points_validate_before_touching_memory():
    """mgr_composite_u8, the ragged pair and mgr_warp_backward: argument errors come back as codes with a message
    (no GPU needed: every call below fails validation or is an empty batch)."""
    import ctypes
    lib = _lib.load()
    assert lib.mgr_composite_u8(0x1000, None, None, None, 2, 3, 8, 8, 0, 0, None) == 1 and b"both NULL" in lib.mgr_last_error()
    assert lib.mgr_composite_u8(0x1000, None, 0x2000, None, 0, 3, 8, 8, 0, 0, None) == 0            # empty batch
    assert lib.mgr_composite_u8(0x1000, None, 0x2000, None, 2, 3, 8, 8, 7, 0, None) == 1            # bad dtype
    lay = (_lib.MgrLayer * 2)(_lib.MgrLayer(0x1000, 4 * 64, 64, 8, 8, 8, 0, 0), _lib.MgrLayer(0x9000, 4 * 16, 16, 4, 4, 4, 2, 2))
    assert lib.mgr_render_forward_ragged(lay, None, 0x2000, None, 2, 2, 8, 8, 0, 0, None) == 2 and b"theta" in lib.mgr_last_error()
    assert lib.mgr_render_forward_ragged(None, 0x3000, 0x2000, None, 2, 2, 8, 8, 0, 0, None) == 1
    assert lib.mgr_render_forward_ragged(lay, 0x3000, 0x2000, None, 2, 1, 8, 8, 0, 0, None) == 2     # a ragged stack needs >= 2 layers
    assert lib.mgr_render_forward_ragged(lay, 0x3000, 0x2000, None, 0, 2, 8, 8, 0, 0, None) == 0     # empty batch
    assert lib.mgr_render_forward_ragged(lay, 0x3000, 0x2000, None, 2, 2, 8, 8, 0, 0, None) == 2     # left offset 2: not a multiple of 4
    assert b"multiples of 4" in lib.mgr_last_error()
    assert lib.mgr_render_backward_ragged(lay, 0x3000, 0x2000, 0x2000, 0x4000, lay, 0x5000, None, 0, 2, 2, 8, 8, 0, 0, 3, None) == 3
    assert b"workspace" in lib.mgr_last_error()
    need = lib.mgr_warp_backward_workspace_bytes(2, 3, 8, 8, _lib.MGR_F32, 3)
    assert need == 2 * 3 * (128 + 4 + 4) + 32 + 2 * 4                                                 # gather path bookkeeping
    assert lib.mgr_warp_backward_workspace_bytes(2, 3, 8, 8, _lib.MGR_BF16, 1) == 2 * 3 * 4 * 64 * 4  # fp32 scatter accumulator dominates
    assert lib.mgr_warp_backward(0x1000, None, 0x3000, 0x2000, 0x4000, 0x5000, None, 0, 2, 3, 8, 8, 0, 0, 3, None) == 3


def test_debug_path_setter_accepts_every_documented_path():
    """include/montage_render.h documents paths 0..4 (4 = the staged stencil kernels instead of the ones on TMA box copies).
    The setter once rejected 4 while its callers ignored the return code, so an A/B run compared a kernel with itself."""
    lib = _lib.load()
    try:
        for path in range(5):
            assert lib.mgr_set_debug_path(path) == 0, path
        assert lib.mgr_set_debug_path(5) == 1 and b"debug path" in lib.mgr_last_error()
        assert lib.mgr_set_debug_path(-1) == 1
    finally:
        lib.mgr_set_debug_path(0)


def test_augment_geom_entry_points_validate():
    lib = _lib.load()
    B, C, H, W, m = 2, 4, 16, 16, (3, 2, 1, 0)
    Hp, Wp = H + m[1] + m[3], W + m[0] + m[2]
    per_plane = (2 * Hp) * (2 * Wp) + (2 * (H + 6)) * (2 * (W + 6)) + Hp * Wp + max(H * 2 * Wp, 2 * (H + 6) * W, 2 * Hp * Wp)
    assert lib.mgr_augment_geom_workspace_bytes(B, C, H, W, *m) == 4 * B * C * per_plane
    assert lib.mgr_augment_geom_workspace_bytes(B, C, H, W, 16, 0, 0, 0) == 0                       # margin > size - 1
    assert lib.mgr_augment_geom_forward(0x1000, 0x2000, 0x3000, None, 0, B, C, H, W, 16, 0, 0, 0, None) == 1
    assert b"margins" in lib.mgr_last_error()
    assert lib.mgr_augment_geom_forward(0x1000, 0x2000, 0x3000, None, 0, B, C, H, W, *m, None) == 3  # workspace too small
    assert lib.mgr_augment_geom_forward(None, 0x2000, 0x3000, None, 0, B, C, H, W, *m, None) == 1
    assert lib.mgr_augment_geom_backward(0x1000, 0x2000, 0x3000, None, 0, 0, C, H, W, *m, None) == 0  # empty batch
