"""CPU restatement of the StyleGAN2 custom ops (test infrastructure -- see oracle/__init__.py).

Follows reference models/stylegan2/op/upfirdn2d.py:159-200 (upfirdn2d_native),
models/stylegan2/op/fused_bias_act_kernel.cu:18-49 (the act/grad table of the CUDA kernel),
models/stylegan2/op/fused_act.py:20-38,86-94 and models/stylegan2/networks.py:236-253,291-298.
All math in the input dtype's fp32/fp64 on CPU with plain torch tensor ops.
"""
import torch
import torch.nn.functional as F


def upfirdn2d_ref(x, kernel, up=1, down=1, pad=(0, 0)):
    """Same signature as the reference's public `upfirdn2d` (upfirdn2d.py:145-156)."""
    return upfirdn2d_ref_full(x, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def upfirdn2d_ref_full(x, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """zero-insert upsample -> pad (negative pads crop) -> TRUE convolution with `kernel` -> decimate.
    (upfirdn2d.py:159-200; the CUDA kernel flips the taps the same way, upfirdn2d_kernel.cu:130-141.)"""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    planes = x.reshape(n * c, h, w)
    # 1. zero insertion: sample i lands on row i*up (upfirdn2d.py:168-170)
    up = planes.new_zeros(n * c, h * up_y, w * up_x)
    up[:, ::up_y, ::up_x] = planes
    # 2. pad with zeros, crop where a pad is negative (:172-180)
    up = F.pad(up, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    up = up[:, max(-py0, 0): up.shape[1] - max(-py1, 0), max(-px0, 0): up.shape[2] - max(-px1, 0)]
    # 3. convolution = correlation with the flipped filter (:185-187)
    taps = torch.flip(kernel.to(up.dtype), [0, 1]).reshape(1, 1, kh, kw)
    full = F.conv2d(up.unsqueeze(1), taps).squeeze(1)
    # 4. decimate (:195)
    out = full[:, ::down_y, ::down_x]
    out_h = (h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (w * up_x + px0 + px1 - kw) // down_x + 1
    assert out.shape[1:] == (out_h, out_w), (out.shape, out_h, out_w)
    return out.reshape(n, c, out_h, out_w)


def fused_bias_act_ref(x, bias, ref, act, grad, alpha, scale):
    """The native op `fused.fused_bias_act` (fused_bias_act.cpp:11-17) element for element:
    bias is broadcast along dim 1 (fused_bias_act_kernel.cu:67-71), then the act*10+grad table (:28-47)."""
    if bias is not None and bias.numel():
        shape = [1, -1] + [1] * (x.dim() - 2)
        x = x + bias.to(x.dtype).reshape(shape)
    if grad == 2:
        y = torch.zeros_like(x)
    elif act == 3:
        gate = x if grad == 0 else ref
        y = torch.where(gate > 0, x, x * alpha)
    elif act == 1:
        y = x
    else:
        raise NotImplementedError(act)
    return y * scale


def fused_leaky_relu_ref(x, bias, negative_slope=0.2, scale=2 ** 0.5):
    """fused_act.py:86-97 with the CUDA semantics (negative_slope honoured; the CPU branch hard-codes 0.2)."""
    return fused_bias_act_ref(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_backward_ref(grad_output, out, negative_slope=0.2, scale=2 ** 0.5):
    """FusedLeakyReLUFunctionBackward.forward (fused_act.py:20-38): grad_input and grad_bias."""
    gx = fused_bias_act_ref(grad_output, None, out, 3, 1, negative_slope, scale)
    dims = [0] + list(range(2, gx.dim()))
    return gx, gx.sum(dims)


def noise_bias_act_ref(x, noise, noise_weight, bias, negative_slope=0.2, scale=2 ** 0.5):
    """NoiseInjection (networks.py:291-298) followed by FusedLeakyReLU (networks.py:346-348)."""
    if noise is not None:
        x = x + noise_weight.to(x.dtype) * noise.to(x.dtype)
    return fused_leaky_relu_ref(x, bias, negative_slope, scale)


def blur_noise_bias_act_ref(x, kernel, pad, noise, noise_weight, bias, negative_slope=0.2, scale=2 ** 0.5,
                            row_scale=None):
    """Blur (networks.py:70-86) -> NoiseInjection -> FusedLeakyReLU, the tail of an upsampling StyledConv."""
    t = upfirdn2d_ref(x, kernel, pad=pad)
    if row_scale is not None:
        t = t * row_scale.reshape(t.shape[0], t.shape[1], 1, 1).to(t.dtype)
    return noise_bias_act_ref(t, noise, noise_weight, bias, negative_slope, scale)


def make_kernel(k):
    """networks.py:17-25: outer product of a 1-D filter, normalised to unit sum."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


def modulated_weight_ref(weight, style, scale, demodulate=True, eps=1e-8):
    """ModulatedConv2d.forward's weight path (networks.py:236-253, run_fp32 / normalize=False branch).
    weight: (1, O, I, k, k); style: (B, I) (already through the modulation EqualLinear).
    Returns (B, O, I, k, k)."""
    b = style.shape[0]
    w = scale * weight * style.reshape(b, 1, -1, 1, 1)
    if demodulate:
        demod = torch.rsqrt(w.pow(2).sum([2, 3, 4]) + eps)
        w = w * demod.reshape(b, -1, 1, 1, 1)
    return w
