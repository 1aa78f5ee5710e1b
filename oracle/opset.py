"""The CPU op namespace (test infrastructure): same entry names as gangealing_b200.opset.cuda_ops(), every
entry a restatement from oracle/*.py.  tests/ and bench.py's CPU-baseline legs build the host-side networks
with `ops=cpu_ops()` to run the reference algorithm end to end on the CPU; the product never does."""
import types

import torch
import torch.nn.functional as F

from . import flow as _flow
from . import perceptual as _perc
from . import sampling as _smp
from . import splat as _splat
from . import stylegan2_ops as _so


def _modulated_weight(weight, style, scale, demodulate=True, transposed=False, eps=1e-8):
    w = _so.modulated_weight_ref(weight, style, scale, demodulate, eps)      # (B, O, I, k, k)
    b, o, i, kh, kw = w.shape
    if transposed:                                                            # networks.py:256-262
        return w.transpose(1, 2).reshape(b * i, o, kh, kw)
    return w.reshape(b * o, i, kh, kw)


def _modulated_conv2d(x, weight, style, scale, demodulate=True, upsample=False, padding=1, eps=1e-8, bias=None, skip=None):
    """The REFERENCE formulation (networks.py:233-282): materialise per-sample filters, grouped convolution.
    Returns (out, None): demodulation is already inside the filters."""
    b, cin, h, w = x.shape
    cout = weight.shape[1]
    wt = _modulated_weight(weight, style, scale, demodulate, transposed=upsample, eps=eps).type(x.dtype)
    xin = x.reshape(1, b * cin, h, w)
    if upsample:
        out = F.conv_transpose2d(xin, wt, padding=0, stride=2, groups=b)
    else:
        out = F.conv2d(xin, wt, padding=padding, groups=b)
    out = out.view(b, cout, out.shape[2], out.shape[3])
    if bias is not None:       # ToRGB epilogue, reference networks.py:400-405
        out = out + bias.type(out.dtype)
    if skip is not None:
        out = out.float() + skip
    return out, None


def _channel_scale(x, s):
    return x * s.reshape(x.shape[0], x.shape[1], 1, 1).to(x.dtype)


def _noise_bias_act(x, noise, noise_weight, bias, negative_slope=0.2, scale=2 ** 0.5, row_scale=None):
    if row_scale is not None:
        x = _channel_scale(x, row_scale)
    return _so.noise_bias_act_ref(x, noise, noise_weight, bias, negative_slope, scale)


def _mipmap_warp(inputs, grid, max_num_levels=8, min_level=0.0, padding_mode="border"):
    out, aux = _smp.mipmap_warp_ref(inputs, grid, max_num_levels, min_level, padding_mode, return_aux=True)
    return out, aux["levels"]


def _bilinear_downsample(x, stride, kernel_horz, kernel_vert):
    return _smp.bilinear_downsample_ref(x, stride)


def _flow_compose(low, mask, identity, base_warp=None, alpha=None, downsample=8):
    return _flow.flow_compose_ref(low, mask, identity, base_warp, alpha, downsample)


_ops = None


def cpu_ops():
    global _ops
    if _ops is None:
        _ops = types.SimpleNamespace(
            name="cpu-oracle",
            upfirdn2d=_so.upfirdn2d_ref,
            fused_leaky_relu=_so.fused_leaky_relu_ref,
            noise_bias_act=_noise_bias_act,
            blur_noise_bias_act=_so.blur_noise_bias_act_ref,
            conv2d=F.conv2d,
            conv_transpose2d=F.conv_transpose2d,
            modulated_weight=_modulated_weight,
            modulated_conv2d=_modulated_conv2d,
            channel_scale=_channel_scale,
            mipmap_warp=_mipmap_warp,
            grid_sample=_smp.warp_ref,
            bilinear_downsample=_bilinear_downsample,
            flow_compose=_flow_compose,
            splat2d=_splat.splat2d_ref,
            feature_distance=_perc.feature_distance_ref,
        )
    return _ops


def fill_parameters(module, seed=0, gain=1.0):
    """Deterministic, construction-order-independent initialisation: every parameter/buffer is drawn from a
    generator seeded by (seed, its qualified name).  Lets the golden script (reference modules) and the tests
    (this repo's modules) hold bit-identical weights without shipping checkpoints."""
    import zlib
    with torch.no_grad():
        for name, t in list(module.named_parameters()) + list(module.named_buffers()):
            if not t.is_floating_point():
                continue
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
            leaf = name.rsplit(".", 1)[-1]
            if leaf in ("kernel", "blur_filter", "kernel_horz", "kernel_vert", "one_hot", "identity_flow"):
                continue  # derived constants
            vals = torch.randn(t.shape, generator=g) * gain
            if leaf == "bias" or name.endswith("noise.weight"):
                vals = vals * 0.1
            t.copy_(vals.to(t.dtype))
    return module


def fill_convs_in_order(module, seed):
    """He-scaled seeded weights for every Conv2d of `module`, in construction order -- the same values for two
    implementations of the same convolution stack whose parameter NAMES differ (reference LPIPS vs this repo's)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    return module
