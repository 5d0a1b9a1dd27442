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


def test_every_public_entry_refuses_cpu_tensors():
    """No CPU fallback anywhere in the product: each public function raises MontageRenderError for CPU tensors instead of
    computing something (the oracle under oracle/ is test infrastructure and is never reached from the package)."""
    from montage_gan_b200 import augment as A, render as mr
    x = torch.zeros(1, 3, 4, 16, 16)
    th = torch.eye(2, 3).expand(1, 3, 2, 3).contiguous()
    layers = [torch.zeros(1, 4, 16, 16), torch.zeros(1, 4, 8, 8)]
    calls = [
        lambda: mr.render(x, th),
        lambda: mr.render(x, None),
        lambda: mr.warp(x, th),
        lambda: mr.alpha_composite_pytorch((x + 1) / 2),
        lambda: mr.alpha_composite((x + 1) / 2),
        lambda: mr.render_ragged(layers, th[:, :2], canvas=(16, 16)),
        lambda: mr.convert_translate_to_2x3(torch.zeros(1, 3, 2)),
        lambda: mr.random_position((x + 1) / 2),
        lambda: A.geometric_warp(torch.zeros(1, 4, 16, 16), torch.eye(3)),
    ]
    for k, call in enumerate(calls):
        with pytest.raises(_lib.MontageRenderError):
            call()


def test_package_never_imports_the_oracle():
    """The product path must not route through oracle/: no module of the package mentions it."""
    import os
    pkg = os.path.dirname(montage_gan_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_stale_library_is_not_loaded_silently(monkeypatch):
    """If the sources changed and the rebuild fails, load() raises instead of loading the old .so (round-1 ADVICE);
    MGR_ALLOW_STALE_LIB=1 opts in, with a warning."""
    import os
    from montage_gan_b200 import build as mbuild
    if not os.path.isfile(_lib.LIB_PATH):
        pytest.skip("library not built")

    def boom(*a, **k):
        raise RuntimeError("nvcc exploded")

    monkeypatch.setattr(mbuild, "build", boom)
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.delenv("MGR_ALLOW_STALE_LIB", raising=False)
    with pytest.raises(_lib.MontageRenderError, match="stale"):
        _lib.load()
    monkeypatch.setenv("MGR_ALLOW_STALE_LIB", "1")
    with pytest.warns(RuntimeWarning, match="STALE"):
        assert _lib.load() is not None
