"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute)."""
import pytest
import torch

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import _lib, render as mr, synth


def test_cpu_tensor_raises_no_fallback():
    x = torch.zeros(1, 2, 4, 8, 8)
    with pytest.raises(_lib.MontageRenderError, match="no CPU path"):
        mr.render(x)
    with pytest.raises(_lib.MontageRenderError):
        mr.alpha_composite_pytorch(x)


@pytest.mark.parametrize("shape", [(2, 4, 8, 8), (1, 2, 3, 8, 8), (1, 0, 4, 8, 8)])
def test_bad_x_shape(shape):
    with pytest.raises(ValueError):
        mr.render(torch.zeros(*shape))


def test_bad_arguments():
    x = torch.zeros(1, 2, 4, 8, 8)
    with pytest.raises(ValueError, match="theta"):
        mr.render(x, torch.zeros(1, 3, 2, 3))
    with pytest.raises(ValueError, match="in_range"):
        mr.render(x, None, in_range="0255")
    with pytest.raises(TypeError):
        mr.render(x.to(torch.int32))
    with pytest.raises(NotImplementedError):
        mr.alpha_composite_pytorch(x, use_premultiplied=True)


def test_synth_is_deterministic_and_in_range():
    for fam in "WSF":
        a = synth.make_layers(2, 5, 16, 12, fam, seed=3)
        b = synth.make_layers(2, 5, 16, 12, fam, seed=3)
        assert torch.equal(a, b) and a.shape == (2, 5, 4, 16, 12)
        assert a.min() >= -1 and a.max() <= 1
    f = synth.make_layers(1, 9, 64, 64, "F", seed=0)
    alpha = (f[:, :, 3] + 1) / 2
    assert (alpha == 0).any() and (alpha == 1).any()
    for fam in "IT0X":
        t = synth.make_theta(2, 5, fam, seed=3)
        assert t.shape == (2, 5, 2, 3) and torch.equal(t, synth.make_theta(2, 5, fam, seed=3))
    t = synth.make_theta(2, 5, "T", seed=3, cover_back=False)
    assert torch.equal(t[..., :2], torch.eye(2).expand(2, 5, 2, 2)) and t[..., 2].abs().max() <= 1
