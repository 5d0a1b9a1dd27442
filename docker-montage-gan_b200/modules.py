"""``nn.Module`` drop-ins for the reference's plug-in seams.

The reference builds its global-GAN pieces by class name (``dnnlib.util.construct_class_by_name``,
``train_aio.py:204`` for the placement net, ``custom/training_loop_aio.py:94-124`` for the
renderer) and passes them to ``MontageGANLoss(pos_estimator=..., renderer=...)``
(``custom/loss_aio.py:199, 232``).  These classes keep those constructors and return types:

* ``AnalyticRenderer(img_resolution, img_channels, img_layers)`` -- signature of
  ``diff_rendering.networks.Renderer*`` (``:7-12, 36-44``): ``[B,L,C,H,W] -> [B,C,H,W]`` in [-1,1].
  No parameters (the learned CNN it replaces only imitates alpha compositing, ``main_united.py:108-127``).
* ``STNv2c`` / ``STNv2b`` -- the placement nets of ``fukuwarai/networks.py:148-258`` with the same
  localisation CNN (plain ``torch.nn`` = cuDNN, out of scope) and the warp done by this library.
  ``fused=True`` returns the *un-warped* layers together with theta and expects the caller to hand
  both to ``FusedRenderer`` so that warp + composite run as one kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import render as _r


class AnalyticRenderer(nn.Module):
    """Exact alpha-over compositor with the learned renderer's call signature."""

    def __init__(self, img_resolution, img_channels=4, img_layers=9, in_range="m11"):
        super().__init__()
        if img_channels != 4:
            raise ValueError("the compositor needs RGBA layers (img_channels == 4)")
        self.img_resolution, self.img_channels, self.img_layers = img_resolution, img_channels, img_layers
        self.in_range = in_range

    def forward(self, x, theta=None):
        """x [B,L,4,H,W] (already placed if theta is None) -> [B,4,H,W]."""
        return _r.render(x, theta, in_range=self.in_range)

    def extra_repr(self):
        return f"img_resolution={self.img_resolution}, img_layers={self.img_layers}, in_range={self.in_range!r}, parameters=0"


class STNv2b(nn.Module):
    """``fukuwarai.networks.STNv2b`` (``:148-226``) with the warp done by libmontage_render (range [0,1])."""
    in_range = "01"

    def __init__(self, img_resolution, img_channels, img_layers, nf1=64, nf2=64, fused=False):
        super().__init__()
        self.img_resolution, self.img_channels, self.img_layers = img_resolution, img_channels, img_layers
        self.fused = fused
        self.localization = nn.Sequential(
            nn.Conv2d(img_channels * img_layers, nf1, kernel_size=7), nn.MaxPool2d(2, stride=2), nn.ReLU(True),
            nn.Conv2d(nf1, nf1 * 2, kernel_size=5), nn.MaxPool2d(2, stride=2), nn.ReLU(True),
            nn.Conv2d(nf1 * 2, nf1 * 4, kernel_size=3), nn.MaxPool2d(2, stride=2), nn.ReLU(True),
            nn.Conv2d(nf1 * 4, nf1 * 6, kernel_size=3), nn.MaxPool2d(2, stride=2), nn.ReLU(True),
            nn.Conv2d(nf1 * 6, nf1 * 8, kernel_size=3), nn.MaxPool2d(2, stride=2), nn.ReLU(True),
        )
        with torch.no_grad():
            self.len_loc = self.localization(
                torch.zeros(1, img_channels * img_layers, img_resolution, img_resolution)).numel()
        self.fc_loc = nn.Sequential(nn.Linear(self.len_loc, nf2), nn.ReLU(True), nn.Linear(nf2, img_layers * 2))
        # identity placement at initialisation (networks.py:201-203)
        self.fc_loc[2].weight.data.zero_()
        self.fc_loc[2].bias.data.zero_()

    def predict_translation(self, x):
        """The localisation CNN + regression head alone (``networks.py:236-245``): ``[B,L,C,H,W] -> [B,L,2]``.  Plain
        ``torch.nn`` on whatever device ``x`` lives on."""
        b, l, c, h, w = x.shape
        feat = self.localization(x.reshape(b, l * c, h, w)).reshape(-1, self.len_loc)
        return self.fc_loc(feat).view(b, l, 2)

    def predict_theta(self, x):
        return _r.convert_translate_to_2x3(self.predict_translation(x))          # [B,L,2,3]

    def forward(self, x):
        """[B,L,C,H,W] -> ([B,L,C,H,W] warped (or un-warped if fused), [B,L,2,3] theta)."""
        theta = self.predict_theta(x)
        if self.fused:
            return x, theta
        return _r.warp(x, theta, in_range=self.in_range), theta


class STNv2c(STNv2b):
    """``fukuwarai.networks.STNv2c`` (``:229-258``): data in [-1,1], the "+1 / grid_sample / -1" warp."""
    in_range = "m11"


class FusedRenderer(AnalyticRenderer):
    """Renderer for the fused path: accepts what a ``fused=True`` placement net returns."""

    def forward(self, x, theta=None):
        return _r.render(x, theta, in_range=self.in_range)
