"""AugmentPipe's geometric execution block (reference: training/augment.py:306-342).

not-gpu: the oracle restatement (oracle/augment_geom.py) against golden vectors produced by the reference pipe itself
         (oracle/make_golden_augment.py): padding margins exactly, sampling matrices and output to fp32 round-off.
gpu:     mgr_augment_geom_* through the C ABI against the golden vectors and the oracle (forward and the gradient
         w.r.t. the images)."""
import os

import numpy as np
import pytest
import torch

import montage_gan_b200  # noqa: F401
from oracle import augment_geom as AG

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "augment_geom_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_oracle_matches_reference_pipe(golden):
    names = [str(n) for n in golden["names"]]
    assert len(names) >= 5
    for n in names:
        x, Gi = torch.from_numpy(golden[f"{n}/images"]), torch.from_numpy(golden[f"{n}/G_inv"])
        H, W = x.shape[2:]
        m = AG.margins(Gi, H, W)
        assert m == tuple(int(v) for v in golden[f"{n}/margins"]), n          # the independent check of the recovered G_inv
        th, (Hs, Ws) = AG.sampling_theta(Gi, H, W, *m)
        assert [Hs, Ws] == [int(v) for v in golden[f"{n}/grid_size"][2:]]
        assert float((th - torch.from_numpy(golden[f"{n}/theta"])).abs().max()) < 1e-6
        y = AG.geometric_warp(x, Gi)
        assert y.shape == x.shape
        assert float((y - torch.from_numpy(golden[f"{n}/out"])).abs().max()) < 5e-5, n


def test_oracle_filters_known_answers():
    f = AG.lowpass_filter()
    assert abs(float(f.sum()) - 1.0) < 1e-6 and f.numel() == 12
    ones = torch.ones(1, 1, 16, 16)
    up = AG.upsample2x(ones, f)
    assert up.shape == (1, 1, 32, 32)
    assert float((up[:, :, 8:24, 8:24] - 1).abs().max()) < 1e-5      # unit DC gain away from the zero-padded border
    dn = AG.downsample2x(torch.ones(1, 1, 44, 44), f, -6)
    assert dn.shape == (1, 1, 16, 16) and float((dn - 1).abs().max()) < 1e-5
    # identity transform: the block is a (slightly low-passed) identity in the interior
    x = torch.rand(1, 2, 32, 32, generator=torch.Generator().manual_seed(1))
    xs = torch.nn.functional.avg_pool2d(x, 5, 1, 2)                     # a smooth image survives the low-pass
    y = AG.geometric_warp(xs, torch.eye(3)[None])
    assert float((y - xs)[:, :, 4:-4, 4:-4].abs().max()) < 2e-2


def test_oracle_index_sampler_equals_grid_sample():
    """oracle.bilinear_sample (the twice-differentiable stand-in) against ATen's grid_sample, value and first gradient."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 17, 23, generator=g, dtype=torch.float64).requires_grad_(True)
    theta = torch.eye(2, 3, dtype=torch.float64).repeat(2, 1, 1) + 0.3 * torch.randn(2, 2, 3, generator=g, dtype=torch.float64)
    grid = torch.nn.functional.affine_grid(theta, [2, 3, 20, 26], align_corners=False)
    a = AG.bilinear_sample(x, grid)
    b = torch.nn.functional.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    assert float((a - b).abs().max()) < 1e-12
    go = torch.randn(a.shape, generator=g, dtype=torch.float64)
    (ga,), (gb,) = torch.autograd.grad(a, x, go), torch.autograd.grad(b, x, go)
    assert float((ga - gb).abs().max()) < 1e-12
    xs = torch.rand(1, 2, 24, 24, generator=g)
    Gi = torch.tensor([[[0.9, 0.2, 1.5], [-0.2, 1.1, -2.0], [0.0, 0.0, 1.0]]])
    assert float((AG.geometric_warp(xs, Gi, double_backward=True) - AG.geometric_warp(xs, Gi)).abs().max()) < 1e-6


def test_host_side_margins_and_matrices_match_reference(golden):
    """The product's host logic (montage_gan_b200.augment: padding margins, the matrices for the sampler) is plain CPU
    arithmetic: it must reproduce what the reference pipe asked of F.pad and affine_grid."""
    from montage_gan_b200 import augment as A
    for n in [str(v) for v in golden["names"]]:
        Gi = torch.from_numpy(golden[f"{n}/G_inv"])
        H, W = golden[f"{n}/images"].shape[2:]
        m = A.padding_margins(Gi, H, W)
        assert m == tuple(int(v) for v in golden[f"{n}/margins"]), n
        th = A.sampling_theta(Gi, H, W, *m)
        assert float((th - torch.from_numpy(golden[f"{n}/theta"])).abs().max()) < 1e-6, n


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_matches_reference_golden(golden):
    from montage_gan_b200 import augment as A
    for n in [str(v) for v in golden["names"]]:
        x, Gi = torch.from_numpy(golden[f"{n}/images"]).cuda(), torch.from_numpy(golden[f"{n}/G_inv"])
        assert A.padding_margins(Gi, x.shape[2], x.shape[3]) == tuple(int(v) for v in golden[f"{n}/margins"]), n
        y = A.geometric_warp(x, Gi)
        assert y.shape == x.shape
        err = float((y.cpu() - torch.from_numpy(golden[f"{n}/out"])).abs().max())
        assert err < 5e-5, (n, err)                                     # fp32 round-off of the sampling coordinates


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 4, 64, 64), (1, 3, 40, 56), (3, 4, 33, 47)])
def test_cuda_forward_and_gradient_match_oracle(shape):
    from montage_gan_b200 import augment as A
    B, C, H, W = shape
    g = torch.Generator().manual_seed(7)
    x = torch.rand(shape, generator=g) * 2 - 1
    ang = (torch.rand(B, generator=g) - 0.5) * 1.5
    sc = 0.8 + 0.5 * torch.rand(B, generator=g)
    Gi = torch.eye(3).repeat(B, 1, 1)
    Gi[:, 0, 0], Gi[:, 0, 1] = sc * torch.cos(ang), -sc * torch.sin(ang) * 1.1
    Gi[:, 1, 0], Gi[:, 1, 1] = sc * torch.sin(ang), sc * torch.cos(ang) * 0.9
    Gi[:, 0, 2], Gi[:, 1, 2] = (torch.rand(B, generator=g) - 0.5) * 0.3 * W, (torch.rand(B, generator=g) - 0.5) * 0.3 * H
    go = torch.randn(shape, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = AG.geometric_warp(xr, Gi)
    (gref,) = torch.autograd.grad(ref, xr, go)
    xd = x.cuda().requires_grad_(True)
    out = A.geometric_warp(xd, Gi)
    (gd,) = torch.autograd.grad(out, xd, go.cuda())
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 5e-5
    assert float((gd.cpu() - gref).abs().max() / gref.abs().max()) < 1e-4


@pytest.mark.gpu
def test_cuda_double_backward_is_the_forward_operator():
    """The block is linear: d/d(grad_out) of <A^T grad_out, v> is A v.  The backward used to detach grad_out, so under
    create_graph=True the result silently had no graph (round-1 ADVICE, high)."""
    from montage_gan_b200 import augment as A
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 2, 4, 40, 48
    Gi = torch.eye(3).repeat(B, 1, 1)
    Gi[:, 0, 1], Gi[:, 1, 0], Gi[:, 0, 2] = 0.2, -0.15, 3.5
    x = (torch.rand(B, C, H, W, generator=g) * 2 - 1).cuda().requires_grad_(True)
    go = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    v = torch.randn(B, C, H, W, generator=g).cuda().requires_grad_(True)
    w = torch.randn(B, C, H, W, generator=g).cuda()
    out = A.geometric_warp(x, Gi)
    (gi,) = torch.autograd.grad(out, x, go, create_graph=True)
    assert gi.requires_grad                                             # the graph through grad_out survives
    (ggo,) = torch.autograd.grad(gi, go, v, create_graph=True)
    assert float((ggo - A.geometric_warp(v.detach(), Gi)).abs().max()) < 1e-5
    # third order: ggo = A v, so its gradient w.r.t. v contracted with w is A^T w -- the first-order backward again
    (third,) = torch.autograd.grad(ggo, v, w)
    (adj,) = torch.autograd.grad(A.geometric_warp(x, Gi), x, w)
    assert float((third - adj).abs().max()) < 1e-5


@pytest.mark.gpu
def test_r1_penalty_through_the_block_matches_oracle():
    """custom/loss_aio.py:327-338 with the augment pipe of :252-254 in between: r1_grads = autograd.grad(logits.sum(),
    images, create_graph=True); the penalty's gradient w.r.t. D's weights flows through the block's double backward
    (reference: grid_sample_gradfix.py:49-88).  D here is a two-layer conv net with a smooth non-linearity."""
    from montage_gan_b200 import augment as A
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    B, C, H, W = 2, 4, 48, 40
    x = torch.rand(B, C, H, W, generator=g) * 2 - 1
    ang = torch.tensor([0.3, -0.5])
    Gi = torch.eye(3).repeat(B, 1, 1)
    Gi[:, 0, 0], Gi[:, 0, 1], Gi[:, 1, 0], Gi[:, 1, 1] = torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang)
    Gi[:, 0, 2], Gi[:, 1, 2] = torch.tensor([2.25, -4.5]), torch.tensor([-1.75, 3.0])
    w1 = torch.randn(8, C, 3, 3, generator=g) * 0.3
    w2 = torch.randn(1, 8, 3, 3, generator=g) * 0.3

    def r1(geom, x_, w1_, w2_):
        x_ = x_.requires_grad_(True)
        w1_, w2_ = w1_.requires_grad_(True), w2_.requires_grad_(True)
        logits = F.conv2d(torch.tanh(F.conv2d(geom(x_, Gi), w1_, padding=1)), w2_, padding=1).mean([1, 2, 3])
        (r1_grads,) = torch.autograd.grad(logits.sum(), x_, create_graph=True, only_inputs=True)
        penalty = r1_grads.square().sum([1, 2, 3]).mean()
        gw = torch.autograd.grad(penalty, [w1_, w2_])
        return penalty.detach().cpu().double(), [t.detach().cpu().double() for t in gw]

    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        pen, gw = r1(A.geometric_warp, x.cuda(), w1.cuda(), w2.cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    pen_ref, gw_ref = r1(lambda im, G: AG.geometric_warp(im, G, double_backward=True), x.double(), w1.double(), w2.double())
    assert float(gw_ref[0].abs().max()) > 0 and float(gw_ref[1].abs().max()) > 0
    assert abs(float(pen - pen_ref)) <= 1e-4 * abs(float(pen_ref))
    for a, r in zip(gw, gw_ref):
        assert float((a - r).abs().max() / r.abs().max()) < 1e-3


@pytest.mark.gpu
def test_cuda_argument_rules():
    from montage_gan_b200 import _lib, augment as A
    x = torch.zeros(1, 4, 16, 16, device="cuda")
    with pytest.raises(ValueError):
        A.geometric_warp(x.half(), torch.eye(3))
    with pytest.raises(ValueError):
        A.geometric_warp(x, torch.eye(3).repeat(2, 1, 1))
    with pytest.raises(_lib.MontageRenderError):
        A.geometric_warp(x.cpu(), torch.eye(3))
    y = A.geometric_warp(x + 0.25, torch.eye(3))                        # identity transform of a constant image
    assert float((y - 0.25).abs().max()) < 1e-5
