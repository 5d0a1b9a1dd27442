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
