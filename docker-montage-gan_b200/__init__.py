"""montage_gan_b200 -- B200-native analytic renderer for MontageGAN's global GAN.

One hot path, nothing else: warp every RGBA layer by its 2x3 placement
(affine_grid + bilinear grid_sample semantics) and alpha-over composite back to front, forward
and backward, as hand-written sm_100a CUDA behind a C ABI (``include/montage_render.h``).
"""
__version__ = "0.1.0"
