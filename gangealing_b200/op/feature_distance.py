"""Perceptual-loss front end in one pass (csrc/lpips.cu): unit-normalise two feature maps over channels, squared
difference, optional per-channel weights, spatial mean.

reference: models/losses/lpips.py:26-28 (`normalize_tensor`), :193-205 (difference, `lins` / channel sum), :226
(`spatial_average`).  `feature_distance(f0, f1, weight=None)` returns (N, 1, 1, 1) like the reference's per-layer `res`.
CUDA tensors only (like every op here: no CPU path in the product).  Channels-last fp32 inputs take the fused kernels
(forward: one read of both maps; backward: one read + one write of both); other CUDA layouts / dtypes evaluate the same
formula with ATen ops on the device."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib


def _composite(f0, f1, weight, eps):
    def unit(f):
        return f / (torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True)) + eps)
    d = (unit(f0) - unit(f1)) ** 2
    if weight is not None:
        d = d * weight.reshape(1, -1, 1, 1)
    return d.sum(dim=1, keepdim=True).mean([2, 3], keepdim=True)


def _supported(f0, f1):
    c = f0.shape[1]
    c4 = c // 4
    ok_c = c % 4 == 0 and c <= 1024 and ((c4 < 32 and c4 & (c4 - 1) == 0) or (c4 >= 32 and c4 % 32 == 0))
    return (f0.is_cuda and f0.dtype == torch.float32 and f1.dtype == torch.float32 and f0.shape == f1.shape and ok_c
            and f0.dim() == 4 and _lib.is_nhwc(f0) and _lib.is_nhwc(f1))


class _FeatureDistance(Function):
    @staticmethod
    def forward(ctx, f0, f1, weight, eps):
        _lib.require_cuda(f0, f1, weight)
        n, c, h, w = f0.shape
        lib = _lib.load()
        wt = weight.detach().float().reshape(-1).contiguous() if weight is not None else None
        out = torch.empty(n, dtype=torch.float32, device=f0.device)
        ws = torch.empty(max(1, lib.gg_feature_distance_workspace(n, c, h * w) // 4), dtype=torch.float32, device=f0.device)
        rc = lib.gg_feature_distance_forward(out.data_ptr(), ws.data_ptr(), f0.data_ptr(), f1.data_ptr(), _lib.ptr(wt),
                                             n, c, h * w, eps, _lib.stream())
        _lib.check(rc, "gg_feature_distance_forward")
        ctx.save_for_backward(f0, f1, wt)
        ctx.eps = eps
        return out.reshape(n, 1, 1, 1)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        f0, f1, wt = ctx.saved_tensors
        need0, need1 = ctx.needs_input_grad[:2]
        n, c, h, w = f0.shape
        g = g.reshape(n).float().contiguous()
        g0 = torch.empty_like(f0) if need0 else None
        g1 = torch.empty_like(f1) if need1 else None
        if need0 or need1:
            rc = _lib.load().gg_feature_distance_backward(_lib.ptr(g0), _lib.ptr(g1), g.data_ptr(), f0.data_ptr(),
                                                          f1.data_ptr(), _lib.ptr(wt), n, c, h * w, ctx.eps, _lib.stream())
            _lib.check(rc, "gg_feature_distance_backward")
        return g0, g1, None, None


def feature_distance(f0, f1, weight=None, eps=1e-10):
    """mean_hw sum_c w_c (f0/|f0| - f1/|f1|)^2 -> (N, 1, 1, 1).  `weight`: (C,) non-trainable `lins` weights or None."""
    _lib.require_cuda(f0, f1, weight)
    if _supported(f0, f1) and (weight is None or not weight.requires_grad):
        return _FeatureDistance.apply(f0, f1, weight, float(eps))
    return _composite(f0, f1, weight, eps)
