"""The drop-ins through the reference's own seams (VERDICT r1 "missing" #4).

* ``dnnlib.util.construct_class_by_name(class_name=..., **stn_kwargs)`` builds the placement net
  (``custom/training_loop_aio.py:283-291``, class name from ``train_aio.py:204``); the renderer is built from
  ``(img_resolution, img_channels, img_layers)`` (``training_loop_aio.py:104-105``) and handed to
  ``MontageGANLoss(pos_estimator=..., renderer=...)`` (``custom/loss_aio.py:199, 232``).
* A checkpoint of the reference's ``fukuwarai.networks.STNv2c`` must load ``strict=True`` into ``modules.STNv2c``.

not-gpu: construction by class name (with the reference's own dnnlib when /root/reference is present, else with the
         same importlib semantics), state_dict compatibility both ways, the localisation CNN on CPU against a golden the
         REAL reference module produced (oracle/make_golden_stn.py).
gpu:     predict_theta and the warp of ``modules.STNv2c`` against that golden; the fused pair against the class swap.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

import montage_gan_b200  # noqa: F401
from montage_gan_b200 import modules as M, synth
from oracle import torch_chain as TC

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stn_golden.npz")
needs_reference = pytest.mark.skipif(not TC.reference_available(), reason="/root/reference not present (GPU box)")


def _construct_class_by_name(*args, class_name=None, **kwargs):
    """``dnnlib/util.py:225-292`` when the reference is here; otherwise its semantics (longest importable module prefix,
    then attribute lookup, then call)."""
    if TC.reference_available():
        TC.load_reference()
        import dnnlib  # type: ignore
        return dnnlib.util.construct_class_by_name(*args, class_name=class_name, **kwargs)
    parts = class_name.split(".")
    for i in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ImportError:
            continue
        for p in parts[i:]:
            obj = getattr(obj, p)
        return obj(*args, **kwargs)
    raise ImportError(class_name)


@pytest.fixture(scope="module")
def stn_golden():
    return np.load(GOLDEN)


def _golden_module(g, **kw):
    res, ch, layers, nf1, nf2 = (int(v) for v in g["cfg"])
    m = M.STNv2c(res, ch, layers, nf1=nf1, nf2=nf2, **kw)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    m.load_state_dict(sd, strict=True)
    return m.eval(), (res, ch, layers)


def test_drop_ins_construct_by_class_name():
    stn_kwargs = dict(img_resolution=128, img_channels=4, img_layers=3)          # training_loop_aio.py:283-287
    pe = _construct_class_by_name(class_name="montage_gan_b200.modules.STNv2c", **stn_kwargs)
    assert isinstance(pe, M.STNv2c) and (pe.img_resolution, pe.img_channels, pe.img_layers) == (128, 4, 3)
    pe = pe.train().requires_grad_(False)                                         # init_module, training_loop_aio.py:247-249
    assert not any(p.requires_grad for p in pe.parameters())
    rd = _construct_class_by_name(class_name="montage_gan_b200.modules.AnalyticRenderer", **stn_kwargs)
    assert isinstance(rd, M.AnalyticRenderer) and len(list(rd.parameters())) == 0 and rd.state_dict() == {}
    fused = _construct_class_by_name(class_name="montage_gan_b200.modules.STNv2c", fused=True, **stn_kwargs)
    assert fused.fused and isinstance(_construct_class_by_name(class_name="montage_gan_b200.modules.FusedRenderer", **stn_kwargs),
                                      M.AnalyticRenderer)
    import copy
    ema = copy.deepcopy(pe).eval()                                                # training_loop_aio.py:292
    assert all(torch.equal(a, b) for a, b in zip(ema.state_dict().values(), pe.state_dict().values()))


@needs_reference
def test_loss_object_accepts_the_drop_ins():
    """``MontageGANLoss`` stores whatever it is given as ``pos_estimator`` / ``renderer`` (``loss_aio.py:231-236``) and
    calls ``self.pos_estimator(fake_layer)`` / ``self.renderer(blchw)``: construct it with the drop-ins."""
    TC.load_reference()
    from custom.loss_aio import MontageGANLoss  # type: ignore
    pe, rd = M.STNv2c(128, 4, 3, nf1=4, nf2=8), M.AnalyticRenderer(128, 4, 3)
    loss = MontageGANLoss(device="cpu", global_d_real_use_renderer=True, renderer_retrain_use_real=False,
                          mapping_network=None, layer_names=[], local_G_list=[], augment_pipe_list=[], local_D_list=[],
                          pos_estimator=pe, renderer=rd, global_D=None)
    assert loss.pos_estimator is pe and loss.renderer is rd


@needs_reference
def test_reference_checkpoint_loads_strict_both_ways():
    _, networks = TC.load_reference()
    torch.manual_seed(0)
    ref = networks.STNv2c(256, 4, 9)                                              # the reference's native configuration
    ours = M.STNv2c(256, 4, 9)
    assert ours.len_loc == ref.len_loc == 12800
    missing = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters())
    ref.load_state_dict(ours.state_dict(), strict=True)                           # and a checkpoint of ours goes back
    for cls_ref, cls_ours in ((networks.STNv2b, M.STNv2b),):
        cls_ours(128, 4, 3, nf1=4, nf2=8).load_state_dict(cls_ref(128, 4, 3, nf1=4, nf2=8).state_dict(), strict=True)
    x = synth.make_layers(1, 9, 256, 256, "S", seed=2)
    with torch.no_grad():
        ref.fc_loc[2].bias.uniform_(-0.5, 0.5)
        ours.load_state_dict(ref.state_dict(), strict=True)
        _, theta = ref.eval()(x)
        tr = ours.eval().predict_translation(x)
    assert torch.equal(theta[..., 2], tr)                                         # same ATen CPU kernels, same weights
    assert torch.equal(theta[..., :2], torch.eye(2).expand(1, 9, 2, 2))


def test_localisation_matches_reference_golden_on_cpu(stn_golden):
    m, (res, ch, layers) = _golden_module(stn_golden)
    x = synth.make_layers(2, layers, res, res, "S", seed=5)
    with torch.no_grad():
        tr = m.predict_translation(x)
    ref_theta = torch.from_numpy(stn_golden["theta"])
    assert float((tr - ref_theta[..., 2]).abs().max()) < 1e-6
    assert float(ref_theta[..., 2].abs().max()) > 0.05                              # not the identity placement


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_placement_net_matches_reference_golden_on_gpu(stn_golden):
    m, (res, ch, layers) = _golden_module(stn_golden)
    m = m.cuda()
    x = synth.make_layers(2, layers, res, res, "S", seed=5).cuda()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            warped, theta = m(x)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    ref_theta = torch.from_numpy(stn_golden["theta"])
    assert theta.shape == (2, layers, 2, 3) and warped.shape == x.shape
    assert float((theta.cpu() - ref_theta).abs().max()) < 2e-5                   # cuDNN vs ATen CPU convolutions
    assert torch.equal(theta.cpu()[..., :2], torch.eye(2).expand(2, layers, 2, 2))
    err = float((warped.cpu()[..., ::4, ::4] - torch.from_numpy(stn_golden["warped_sub4"])).abs().max())
    assert err < 2e-4, err       # |d warped / d shift| <= 64 px * 2 per unit theta: 2e-5 of theta -> ~1e-4 at edges


@pytest.mark.gpu
def test_fused_pair_equals_class_swap(stn_golden):
    """``STNv2c(fused=True)`` + ``FusedRenderer(x, theta)`` == ``STNv2c`` + ``AnalyticRenderer(warped)``."""
    plain, (res, ch, layers) = _golden_module(stn_golden)
    fused, _ = _golden_module(stn_golden, fused=True)
    plain, fused = plain.cuda(), fused.cuda()
    x = synth.make_layers(2, layers, res, res, "S", seed=5).cuda()
    with torch.no_grad():
        warped, theta = plain(x)
        a = M.AnalyticRenderer(res, ch, layers)(warped)
        x2, theta2 = fused(x)
        b = M.FusedRenderer(res, ch, layers)(x2, theta2)
    assert torch.equal(theta, theta2) and x2 is x
    assert float((a - b).abs().max()) < 1e-5
