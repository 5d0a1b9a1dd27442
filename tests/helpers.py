"""Shared error metrics and the three-way parity rule (SURVEY.md 8d "Parity reporting")."""
import numpy as np

FWD_TOL = 1e-5      # north-star: forward max-abs (fp32)
GRAD_TOL = 1e-4     # north-star: gradients, relative to max|ref|


def max_abs(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0


def rel_err(a, ref):
    """max|a-ref| / max|ref| over entries where ref is finite."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    m = np.isfinite(ref)
    if not m.any():
        return 0.0
    scale = max(float(np.abs(ref[m]).max()), 1e-30)
    return float(np.abs(np.where(m, a - ref, 0.0)).max()) / scale


def three_way(new, ref32, ref64, tol, metric):
    """Returns (ok, dict).  Pass if new is within tol of the fp32 reference, OR within
    max(tol, 2*floor) of the fp64 reference where floor = err(ref32, ref64) is the reference's own
    fp32 conditioning on these inputs."""
    e32 = metric(new, ref32)
    e64 = metric(new, ref64)
    floor = metric(ref32, ref64)
    ok = (e32 <= tol) or (e64 <= max(tol, 2 * floor))
    return ok, dict(new_vs_ref32=e32, new_vs_ref64=e64, ref32_vs_ref64=floor, tol=tol)
