from .networks import (Blur, ConvLayer, Downsample, EqualConv2d, EqualLinear, FusedLeakyReLU, Generator, ModulatedConv2d,
                       NoiseInjection, PixelNorm, ResBlock, ScaledLeakyReLU, StyledConv, ToRGB, Upsample, make_kernel)
