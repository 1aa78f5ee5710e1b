"""splat2d -- drop-in for reference utils/splat2d_cuda/functional.py:31-64 on sm_100a.

`splat2d(input, coordinates, values, sigma, soft_normalize=False)`: same argument checks (as RuntimeError /
AssertionError messages), CUDA + float32 only, forward only (the reference's backward raises too).
"""
import torch
import torch.autograd as ag

from .. import _lib

__all__ = ["splat2d"]


class Splat2DFunction(ag.Function):
    @staticmethod
    def forward(ctx, input, coordinates, values, sigma, soft_normalize=False):
        assert coordinates.dtype == torch.float32 and values.dtype == torch.float32, \
            "Splat2D only takes float coordinates and values, got {} and {} instead.".format(coordinates.type(), values.type())
        assert coordinates.size(0) == values.size(0) and coordinates.size(1) == values.size(1), \
            "coordinates should be size (N, num_points, 2) and values should be size (N, num_points, *), got {} and {} instead.".format(
                coordinates.shape, values.shape)
        assert input.size(0) == coordinates.size(0) and input.dim() == 4, \
            "input should be of size (N, *, H, W), got {} instead".format(input.shape)
        assert sigma.size(0) == input.size(0), "sigma should be a tensor of size (N,)"
        if not coordinates.is_cuda:
            raise NotImplementedError("Splat2D currently only has support for GPU (cuda).")
        _lib.require_cuda(input, values, sigma)
        if input.dtype != torch.float32 or sigma.dtype != torch.float32:
            raise RuntimeError("splat2d: input and sigma must be float32")
        n, c, h, w = input.shape
        if values.dim() != 3 or values.size(2) != c or coordinates.size(2) != 2:
            raise RuntimeError("splat2d: values must be (N, P, %d) and coordinates (N, P, 2)" % c)
        input, coordinates, values, sigma = [t.contiguous() for t in (input, coordinates, values, sigma)]
        lib = _lib.load()
        out = torch.empty_like(input)
        ws = torch.empty(max(1, lib.gg_splat2d_workspace(n, c, h, w) // 4), dtype=torch.float32, device=input.device)
        rc = lib.gg_splat2d_forward(out.data_ptr(), ws.data_ptr(), input.data_ptr(), coordinates.data_ptr(),
                                    values.data_ptr(), sigma.data_ptr(), n, coordinates.size(1), c, h, w,
                                    1 if soft_normalize else 0, _lib.stream())
        _lib.check(rc, "gg_splat2d_forward")
        return out

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError


splat2d = Splat2DFunction.apply
