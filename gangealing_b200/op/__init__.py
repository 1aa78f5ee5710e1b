"""Op-level API of the StyleGAN2 part of the hot path (mirror of reference models/stylegan2/op/__init__.py)."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d
from . import conv2d_gradfix
from .styled_tail import noise_bias_act, blur_noise_bias_act

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "conv2d_gradfix", "noise_bias_act",
           "blur_noise_bias_act"]
