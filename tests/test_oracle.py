"""CPU tests: the oracle restatement against the golden vectors produced by the REAL reference
(oracle/make_golden.py), against torch's ATen CPU kernels, and -- when /root/reference is present
(build container) -- against the reference modules themselves."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restatement as R
from oracle import torch_chain as TC
from helpers import max_abs, rel_err


def _case(golden, name):
    theta = golden[f"{name}/theta"] if f"{name}/theta" in golden.files else None
    return golden[f"{name}/x"], theta, golden[f"{name}/grad_out"], str(golden[f"{name}/in_range"])


def test_golden_has_all_cases(golden):
    names = list(golden["cases"])
    assert len(names) >= 10
    for n in names:
        assert f"{n}/ref64/out" in golden.files and f"{n}/ref32/grad_x" in golden.files


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 16, 31, 32, 96, 128, 160, 255, 256, 512, 1024])
def test_base_coords_bitwise_vs_aten(n):
    ref = F.affine_grid(torch.eye(2, 3)[None], (1, 1, n, n), align_corners=False)[0, 0, :, 0].numpy()
    assert np.array_equal(R.base_coords(n, np.float32), ref)
    exact = (2 * np.arange(n) + 1) / n - 1
    assert np.abs(R.base_coords(n, np.float64) - exact).max() < 1e-15


def test_affine_grid_bitwise_vs_aten():
    g = torch.Generator().manual_seed(3)
    theta = torch.eye(2, 3) + 0.25 * torch.randn(5, 2, 3, generator=g)
    ref = F.affine_grid(theta, (5, 1, 40, 24), align_corners=False).numpy()
    gx, gy = R.affine_grid(theta.numpy(), 40, 24, np.float32)
    # bitwise on the CPUs seen so far; allow 1 ulp so a different BLAS kernel cannot break CI
    assert np.abs(gx - ref[..., 0]).max() <= 2.4e-7 and np.abs(gy - ref[..., 1]).max() <= 2.4e-7


def test_grid_sample_vs_aten():
    g = torch.Generator().manual_seed(4)
    theta = torch.eye(2, 3) + 0.3 * torch.randn(4, 2, 3, generator=g)
    img = torch.rand(4, 4, 20, 28, generator=g)
    grid = F.affine_grid(theta, img.shape, align_corners=False)
    ref = F.grid_sample(img, grid, align_corners=False).numpy()
    gx, gy = R.affine_grid(theta.numpy(), 20, 28, np.float32)
    out, _ = R.grid_sample_fwd(img.numpy(), gx, gy)
    assert max_abs(out, ref) < 5e-6
    ref64 = F.grid_sample(img.double(), F.affine_grid(theta.double(), img.shape, align_corners=False),
                          align_corners=False).numpy()
    gx, gy = R.affine_grid(theta.numpy(), 20, 28, np.float64)
    out64, _ = R.grid_sample_fwd(img.double().numpy(), gx, gy)
    assert max_abs(out64, ref64) < 1e-13


def test_restatement_matches_golden_fp64(golden):
    for name in golden["cases"]:
        x, theta, go, in_range = _case(golden, name)
        r = R.render_fwd_bwd(x, theta, go, in_range, np.float64)
        assert max_abs(r["out"], golden[f"{name}/ref64/out"]) < 1e-12, name
        ref_gx = golden[f"{name}/ref64/grad_x"]
        assert rel_err(r["grad_x"], ref_gx) < 1e-11, name
        if theta is not None:
            ref_gt = golden[f"{name}/ref64/grad_theta"]
            if np.isfinite(ref_gt).all():
                assert rel_err(r["grad_theta"], ref_gt) < 1e-10, name


def test_restatement_matches_golden_fp32(golden):
    for name in golden["cases"]:
        x, theta, go, in_range = _case(golden, name)
        r = R.render_fwd_bwd(x, theta, go, in_range, np.float32)
        assert max_abs(r["out"], golden[f"{name}/ref32/out"]) < 5e-6, name
        assert rel_err(r["grad_x"], golden[f"{name}/ref32/grad_x"]) < 2e-5, name


def test_reference_nan_where_back_layers_are_transparent(golden):
    """SURVEY.md finding 3: the reference's backward is NaN wherever an ``a_over_b`` step divides
    by a zero canvas alpha -- i.e. where layers 0 and 1 are both fully transparent (the first
    step already yields 0/0), a superset of "nothing covers the pixel".  The restatement gives
    the finite closed-form gradient there, defines 0 where the final alpha is 0, and reports
    those pixels."""
    name = "composite_only_sparse"
    x, theta, go, in_range = _case(golden, name)
    r = R.render_fwd_bwd(x, theta, go, in_range, np.float64)
    ref = golden[f"{name}/ref64/grad_x"]
    nan_px = ~np.isfinite(ref).all(axis=(1, 2))       # [B,H,W]
    a = (x[:, :, 3].astype(np.float64) + 1) / 2
    assert nan_px.any()
    assert np.array_equal(nan_px, (a[:, 0] == 0) & (a[:, 1] == 0))
    assert np.array_equal(r["nan_mask"], (a == 0).all(axis=1))
    assert not (r["nan_mask"] & ~nan_px).any()
    assert np.isfinite(r["grad_x"]).all()
    assert np.all(r["grad_x"][np.broadcast_to(r["nan_mask"][:, None, None], ref.shape)] == 0)


def test_known_answers(golden):
    # layer 0 is the back: opaque green over opaque red is green (image_utils.py:142-146)
    out = R.alpha_composite(golden["ka/order/in"])
    assert np.array_equal(out, golden["ka/order/out"])
    assert np.allclose(out[0, :, 0, 0], [0, 1, 0, 1])
    # fully transparent stack -> exactly 0 (nan_to_num of 0/0, image_utils.py:132)
    out = R.alpha_composite(golden["ka/transparent/in"])
    assert np.array_equal(out, golden["ka/transparent/out"]) and not out.any()
    # three half-transparent layers (image_utils.py:413-420)
    assert max_abs(R.alpha_composite(golden["ka/half/in"]), golden["ka/half/out"]) < 1e-7
    # translation -> theta (image_utils.py:316-335)
    assert np.array_equal(R.convert_translate_to_2x3(golden["ka/translate2x3/in"]), golden["ka/translate2x3/out"])
    # +tx moves content LEFT (image_utils.py:23-28)
    w, _ = R.warp_fwd(golden["ka/shift/in"], golden["ka/shift/theta"], "01", np.float32)
    assert max_abs(w[0], golden["ka/shift/out"]) < 1e-6
    assert w[0, 0, 0, 4].argmax() == 3          # impulse moved from column 5 to column 3 (0.5 * W/2 = 2 px)


def test_single_layer_is_identity(golden):
    x = golden["single_layer/x"]
    out = R.render_fwd(x, None, "m11", np.float32)
    assert max_abs(out, x[:, 0]) < 1e-6


def test_port_chain_matches_golden(golden):
    """The travelling torch port (CPU arm of bench.py) reproduces the real reference's outputs."""
    for name in golden["cases"]:
        x, theta, go, in_range = _case(golden, name)
        xt, got = torch.from_numpy(x), torch.from_numpy(go)
        tt = None if theta is None else torch.from_numpy(theta)
        for tag, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
            r = TC.fwd_bwd(TC.port_chain, xt, tt, got, in_range, dt)
            tol = 1e-12 if dt == torch.float64 else 2e-6
            assert max_abs(r["out"].numpy(), golden[f"{name}/{tag}/out"]) <= tol, (name, tag)
            assert rel_err(r["grad_x"].numpy(), golden[f"{name}/{tag}/grad_x"]) <= max(tol, 1e-5 if dt == torch.float32 else 0), (name, tag)


def test_composite_jvp_is_derivative():
    g = np.random.default_rng(0)
    z = g.random((2, 5, 4, 6, 6)) * 0.9 + 0.05
    dz = g.standard_normal(z.shape)
    eps = 1e-6
    num = (R.alpha_composite(z + eps * dz) - R.alpha_composite(z - eps * dz)) / (2 * eps)
    assert max_abs(R.composite_jvp(z, dz), num) < 1e-7


@pytest.mark.skipif(not TC.reference_available(), reason="/root/reference not present (GPU box)")
def test_restatement_vs_live_reference():
    from montage_gan_b200 import synth
    for lf, tf in (("W", "I"), ("S", "T"), ("W", "X")):
        x = synth.make_layers(2, 4, 24, 20, lf, seed=7)
        th = synth.make_theta(2, 4, tf, seed=7)
        go = synth.make_grad_out(2, 24, 20, seed=7)
        ref = TC.fwd_bwd(TC.reference_chain, x, th, go, "m11", torch.float64)
        port = TC.fwd_bwd(TC.port_chain, x, th, go, "m11", torch.float64)
        r = R.render_fwd_bwd(x.numpy(), th.numpy(), go.numpy(), "m11", np.float64)
        assert torch.equal(ref["out"], port["out"]) and torch.equal(ref["grad_x"], port["grad_x"])
        assert max_abs(r["out"], ref["out"].numpy()) < 1e-13
        assert rel_err(r["grad_x"], ref["grad_x"].numpy()) < 1e-12
        assert rel_err(r["grad_theta"], ref["grad_theta"].numpy()) < 1e-11
