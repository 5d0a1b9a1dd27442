"""End-to-end path with host buffers (what ``bench.py`` reports as ``e2e``): pinned CPU tensors in,
pinned CPU tensors out, chunked H2D / kernels / D2H pipeline inside ``mgr_render_fwd_bwd_host``."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .render import _DTYPES, _RANGES


class HostRenderer:
    """Owns the device workspace and the pinned result buffers for one problem shape."""

    def __init__(self, B, L, H, W, dtype=torch.float32, chunk_B=8, device="cuda:0", in_range="m11"):
        self.lib = _lib.load()
        self.shape = (B, L, H, W)
        self.dtype, self.device, self.in_range = dtype, torch.device(device), in_range
        self.chunk_B = max(1, min(chunk_B, B))
        self.ws_bytes = self.lib.mgr_render_host_workspace_bytes(self.chunk_B, L, H, W, _DTYPES[dtype])
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.out = torch.empty((B, 4, H, W), dtype=dtype).pin_memory()
        self.grad_x = torch.empty((B, L, 4, H, W), dtype=dtype).pin_memory()
        self.grad_theta = torch.empty((B, L, 2, 3), dtype=torch.float32).pin_memory()

    def fwd_bwd(self, x, theta, grad_out, sync=True):
        """x [B,L,4,H,W], theta [B,L,2,3] fp32, grad_out [B,4,H,W]: contiguous CPU tensors (pinned for
        asynchronous copies).  Returns (out, grad_x, grad_theta) pinned CPU tensors owned by this object."""
        B, L, H, W = self.shape
        for t, shp, dt in ((x, (B, L, 4, H, W), self.dtype), (theta, (B, L, 2, 3), torch.float32), (grad_out, (B, 4, H, W), self.dtype)):
            if t.is_cuda or tuple(t.shape) != shp or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"expected a contiguous CPU tensor {shp} {dt}, got {tuple(t.shape)} {t.dtype} cuda={t.is_cuda}")
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.mgr_render_fwd_bwd_host(P(x), P(theta), P(grad_out), P(self.out), P(self.grad_x), P(self.grad_theta),
                                                  P(self.ws), self.ws_bytes, self.chunk_B, B, L, H, W, _DTYPES[self.dtype],
                                                  _RANGES[self.in_range], ctypes.c_void_p(stream.cuda_stream))
            _lib.check(rc, "mgr_render_fwd_bwd_host")
            if sync:
                stream.synchronize()
        return self.out, self.grad_x, self.grad_theta
