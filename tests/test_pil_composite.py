"""The non-differentiable 8-bit composite (reference: custom_utils/image_utils.py:74-96, Pillow arithmetic).
Integer work: every comparison here is bit-exact.

not-gpu: the oracle restatement against the golden vectors made by the real reference (oracle/make_golden_pil.py),
         against Pillow itself when it is importable, and known answers.
gpu:     mgr_composite_u8 through the C ABI against the golden vectors and against the oracle on seeded inputs
         (dtypes, strided views, odd shapes, both ranges), plus size-independent properties at config-2 size."""
import os

import numpy as np
import pytest
import torch

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import synth
from oracle import restatement as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pil_composite_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_matches_reference_golden(golden):
    names = [str(n) for n in golden["names"]]
    assert {"smooth_L7", "sparse_L9", "all_alpha_pairs", "single_layer", "range01", "unbatched"} <= set(names)
    for n in names:
        out, u8 = R.pil_alpha_composite(golden[f"{n}/x"], str(golden[f"{n}/in_range"]))
        assert np.array_equal(out, golden[f"{n}/out"]), n
        assert np.array_equal(out, u8.astype(np.float32) / np.float32(255)), n


def test_oracle_over_matches_pillow_when_available():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for it in range(8):
        dst = rng.integers(0, 256, (48, 40, 4), dtype=np.uint8)
        src = rng.integers(0, 256, (48, 40, 4), dtype=np.uint8)
        if it % 2:
            src[..., 3] = rng.choice([0, 1, 254, 255], size=src.shape[:2])
            dst[..., 3] = rng.choice([0, 1, 128, 255], size=dst.shape[:2])
        a = Image.fromarray(dst, "RGBA")
        a.alpha_composite(Image.fromarray(src, "RGBA"))
        assert np.array_equal(np.asarray(a), R.pil_over(dst, src))


def test_oracle_known_answers():
    red = np.array([255, 0, 0, 255], np.uint8)
    green = np.array([0, 255, 0, 255], np.uint8)
    clear = np.array([9, 9, 9, 0], np.uint8)
    half = np.array([0, 0, 255, 128], np.uint8)
    assert np.array_equal(R.pil_over(red, green), green)                 # opaque front wins (layer order, image_utils.py:85-87)
    assert np.array_equal(R.pil_over(red, clear), red)                   # transparent source copies the canvas
    assert np.array_equal(R.pil_over(clear, half), half)                 # over a transparent canvas: the source itself
    assert np.array_equal(R.pil_over(red, half), np.array([127, 0, 128, 255], np.uint8))
    # byte conversion truncates: 0.999 * 255 = 254.7 -> 254; exactly 1.0 -> 255
    assert R.pil_to_byte(np.float32(0.999)) == 254 and R.pil_to_byte(np.float32(1.0)) == 255
    assert R.pil_to_byte(np.float32(-0.2)) == 0 and R.pil_to_byte(np.float32(1.7)) == 255


# ------------------------------------------------------------------------------------------------ GPU
def _cuda(x, in_range="01", dtype=torch.float32, return_bytes=True):
    from montage_gan_b200 import render as mr
    xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))
    out, u8 = mr.alpha_composite(xt.to("cuda", dtype), in_range=in_range, return_bytes=True)
    return out.cpu().numpy(), u8.cpu().numpy()


@pytest.mark.gpu
def test_cuda_matches_reference_golden(golden):
    for n in [str(v) for v in golden["names"]]:
        out, u8 = _cuda(golden[f"{n}/x"], str(golden[f"{n}/in_range"]))
        assert out.dtype == np.float32 and out.shape == golden[f"{n}/out"].shape
        assert np.array_equal(out, golden[f"{n}/out"]), n
        assert np.array_equal(out, u8.astype(np.float32) / np.float32(255)), n


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 7, 64, 64), (1, 9, 40, 36), (3, 2, 17, 23), (2, 1, 8, 12), (1, 32, 16, 16)])
@pytest.mark.parametrize("family", ["S", "W", "F"])
@pytest.mark.parametrize("in_range", ["m11", "01"])
def test_cuda_matches_oracle(shape, family, in_range):
    B, L, H, W = shape
    x = synth.make_layers(B, L, H, W, family, seed=7)
    if in_range == "01":
        x = (x + 1) / 2
    ref, ref8 = R.pil_alpha_composite(x.numpy(), in_range)
    out, u8 = _cuda(x, in_range)
    assert np.array_equal(u8, ref8)
    assert np.array_equal(out, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_cuda_16bit_storage_matches_oracle_on_rounded_inputs(dtype):
    x = synth.make_layers(2, 7, 48, 48, "S", seed=11).to(dtype)
    ref, ref8 = R.pil_alpha_composite(x.float().numpy(), "m11")
    out, u8 = _cuda(x, "m11", dtype)
    assert np.array_equal(u8, ref8) and np.array_equal(out, ref)


@pytest.mark.gpu
def test_cuda_strided_views_and_unbatched():
    from montage_gan_b200 import render as mr
    big = synth.make_layers(3, 9, 40, 48, "W", seed=3).cuda()
    for view in (big[:, ::2], big[1:, 2:7, :, 4:36, 8:40], big[:, :, :, :, 1:33], big[:, :, :, ::2, ::2]):
        ref, ref8 = R.pil_alpha_composite(view.cpu().numpy(), "m11")
        out, u8 = mr.alpha_composite(view, in_range="m11", return_bytes=True)
        assert np.array_equal(u8.cpu().numpy(), ref8) and np.array_equal(out.cpu().numpy(), ref)
    one = big[0]
    out = mr.alpha_composite(one, in_range="m11")
    assert out.shape == (4, 40, 48)
    assert np.array_equal(out.cpu().numpy(), R.pil_alpha_composite(one.cpu().numpy(), "m11")[0])
    with pytest.raises(montage_gan_b200._lib.MontageRenderError):
        mr.alpha_composite(big.cpu())                                     # no CPU path


@pytest.mark.gpu
def test_cuda_properties_at_config2_size():
    """B=64, L=7, 256x256 bf16 (too large for the numpy oracle to be worth it): properties the arithmetic guarantees."""
    from montage_gan_b200 import render as mr
    x = synth.make_layers(8, 7, 256, 256, "F", seed=5).repeat(8, 1, 1, 1, 1).to("cuda", torch.bfloat16)
    out, u8 = mr.alpha_composite(x, in_range="m11", return_bytes=True)
    # (1) a stack of transparent layers in front changes nothing (transparent source copies the canvas)
    clear = torch.full_like(x[:, :2], -1.0)
    out2, u82 = mr.alpha_composite(torch.cat([x, clear], 1), in_range="m11", return_bytes=True)
    assert torch.equal(u8, u82) and torch.equal(out, out2)
    # (2) an opaque layer in front replaces everything behind it by its own bytes
    front = x[:, 3:4].clone()
    front[:, :, 3] = 1.0
    _, u83 = mr.alpha_composite(torch.cat([x, front], 1), in_range="m11", return_bytes=True)
    _, only = mr.alpha_composite(front, in_range="m11", return_bytes=True)
    assert torch.equal(u83, only)
    # (3) samples are independent: any sub-batch gives the same bytes
    _, part = mr.alpha_composite(x[5:9], in_range="m11", return_bytes=True)
    assert torch.equal(part, u8[5:9])
    # (4) float output is exactly byte / 255 (IEEE division, as ToTensor does on the CPU)
    assert np.array_equal(out[:4].cpu().numpy(), u8[:4].cpu().numpy().astype(np.float32) / np.float32(255))
