"""The reference's render chain on CPU with torch: (a) through the REAL reference modules when
/root/reference exists (build container only), (b) as a self-contained port that travels.

TEST INFRASTRUCTURE ONLY -- see ``oracle/restatement.py`` for the rule.  The product never
imports this.  Used for: pinning ``oracle/restatement.py`` (tests, ``make_golden.py``), and as
the timed CPU arm of ``bench.py`` (``cpu_baseline`` / ``--impl reference``), because it executes
the same ATen CPU kernels and the same per-sample / per-layer Python loops the reference does.

Chain (paths relative to /root/reference/montage_gan), SURVEY.md 3.2:
    x2   = x.view(B*L,4,H,W)                                         fukuwarai/networks.py:250
    grid = F.affine_grid(theta.view(-1,2,3), x2.size(), False)        :251
    w    = F.grid_sample(x2 + 1, grid, align_corners=False) - 1       :253-255
    out  = normalize_minus11(alpha_composite_pytorch(normalize_zero1(w)))   custom/loss_aio.py:251
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("MONTAGE_REFERENCE_ROOT", "/root/reference/montage_gan")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "custom_utils", "image_utils.py"))


def load_reference():
    """Import the real reference modules (read-only).  Returns (image_utils, fukuwarai.networks)."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from custom_utils import image_utils  # type: ignore
    from fukuwarai import networks  # type: ignore
    return image_utils, networks


def reference_chain(x: torch.Tensor, theta, in_range: str = "m11") -> torch.Tensor:
    """SURVEY.md 3.2 evaluated with the reference's own functions (differentiable)."""
    iu, _ = load_reference()
    B, L, C, H, W = x.shape
    w = x
    if theta is not None:
        x2 = x.reshape(-1, C, H, W)
        grid = F.affine_grid(theta.reshape(-1, 2, 3), x2.size(), align_corners=False)
        if in_range == "m11":
            w = (F.grid_sample(x2 + 1, grid, align_corners=False) - 1).view(B, L, C, H, W)
        else:
            w = F.grid_sample(x2, grid, align_corners=False).view(B, L, C, H, W)
    if in_range == "m11":
        return iu.normalize_minus11(iu.alpha_composite_pytorch(iu.normalize_zero1(w)))
    return iu.alpha_composite_pytorch(w)


# ---- self-contained port (same ATen calls, same loop structure) ------------------------------

def _a_over_b(chw1, chw2):
    # custom_utils/image_utils.py:128-133
    color1, alpha1 = chw1[:3], chw1[3:]
    color2, alpha2 = chw2[:3], chw2[3:]
    alpha_out = alpha1 + alpha2 * (1 - alpha1)
    color_out = torch.nan_to_num((color1 * alpha1 + color2 * alpha2 * (1 - alpha1)) / alpha_out)
    return torch.cat([color_out, alpha_out])


def alpha_composite_port(blchw: torch.Tensor) -> torch.Tensor:
    # custom_utils/image_utils.py:142-146 (process) and :163 (stack over the batch)
    outs = []
    for lchw in blchw:
        canvas = lchw[0]
        for chw in lchw[1:]:
            canvas = _a_over_b(chw, canvas)
        outs.append(canvas)
    return torch.stack(outs)


def port_chain(x: torch.Tensor, theta, in_range: str = "m11") -> torch.Tensor:
    """Same chain as ``reference_chain`` without importing /root/reference (it does not exist on
    the GPU box).  Checked equal (bitwise, fp32 and fp64) to ``reference_chain`` in tests."""
    B, L, C, H, W = x.shape
    w = x
    if theta is not None:
        x2 = x.reshape(-1, C, H, W)
        grid = F.affine_grid(theta.reshape(-1, 2, 3), x2.size(), align_corners=False)
        if in_range == "m11":
            w = (F.grid_sample(x2 + 1, grid, align_corners=False) - 1).view(B, L, C, H, W)
        else:
            w = F.grid_sample(x2, grid, align_corners=False).view(B, L, C, H, W)
    if in_range == "m11":
        return alpha_composite_port((w + 1.) / 2.) * 2. - 1.
    return alpha_composite_port(w)


def fwd_bwd(chain, x, theta, grad_out, in_range="m11", dtype=torch.float32):
    """Run ``chain`` forward + autograd backward at ``dtype`` on CPU.  Returns dict of tensors
    (out, grad_x, grad_theta-or-None), all detached."""
    xd = x.detach().to("cpu", dtype).clone().requires_grad_(True)
    td = None if theta is None else theta.detach().to("cpu", dtype).clone().requires_grad_(True)
    out = chain(xd, td, in_range)
    ins = [xd] if td is None else [xd, td]
    grads = torch.autograd.grad(out, ins, grad_out.detach().to("cpu", dtype))
    return dict(out=out.detach(), grad_x=grads[0], grad_theta=None if td is None else grads[1].view(theta.shape))
