"""Modulated-convolution weight path: modulate -> demodulate -> (transpose for the up-conv) in fused kernels.

reference: models/stylegan2/networks.py:233-253 (`weight = scale*W*style`, `demod = rsqrt(sum w^2 + 1e-8)`,
`weight *= demod`, reshape / transpose for conv_transpose2d :255-262): ~6 ATen launches that materialise three
(B, O, I, k, k) temporaries per layer.  `modulated_weight` is the op-level entry used by ModulatedConv2d.
"""
import torch


def modulated_weight_composite(weight, style, scale, demodulate=True, transposed=False, eps=1e-8):
    """Plain tensor-op formulation (differentiable w.r.t. everything); used when the filter bank itself needs a
    gradient (never the case for GANgealing's frozen generator) until the fused kernels take over that case too."""
    b = style.shape[0]
    _, o, i, kh, kw = weight.shape
    w = (scale * weight) * style.reshape(b, 1, i, 1, 1)
    if demodulate:
        w = w * torch.rsqrt(w.pow(2).sum([2, 3, 4]) + eps).reshape(b, o, 1, 1, 1)
    if transposed:
        return w.transpose(1, 2).reshape(b * i, o, kh, kw)
    return w.reshape(b * o, i, kh, kw)


def modulated_weight(weight, style, scale, demodulate=True, transposed=False, eps=1e-8):
    """weight (1, O, I, k, k), style (B, I) -> per-sample filters for the grouped convolution:
    (B*O, I, k, k), or (B*I, O, k, k) when `transposed` (the layout conv_transpose2d(groups=B) wants)."""
    return modulated_weight_composite(weight, style, scale, demodulate, transposed, eps)
