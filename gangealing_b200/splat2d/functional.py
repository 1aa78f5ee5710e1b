"""splat2d -- drop-in for reference utils/splat2d_cuda/functional.py:31-64 on sm_100a.

`splat2d(input, coordinates, values, sigma, soft_normalize=False)`: same argument checks (as RuntimeError /
AssertionError messages), CUDA + float32 only, forward only (the reference's backward raises too).
"""
import torch
import torch.autograd as ag

from .. import _lib

__all__ = ["splat2d", "splat2d_lookup", "nn_argmin"]


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


def splat2d_lookup(input, grid, query, values, sigma, res, out_res, soft_normalize=False):
    """`uncongeal_points`' lookup fused into the splat (csrc/splat.cu LOOKUP):
        points = unnormalize(F.grid_sample(grid as image, query, 'border'), res, out_res)      spatial_transformer.py:141-157,621-623
        out    = splat2d(input, points, values, sigma, soft_normalize)                         functional.py:31-64
    in ONE scatter pass: the sampling grid is read where the points are loaded.  grid: (N, Hg, Wg, 2) sampling grid of the STN;
    query: (N, P, 2) normalised congealed-frame coordinates; values: (N, P, C <= 3).  -> (out, points (N, P, 2) pixels)."""
    _lib.require_cuda(input, grid, query, values, sigma)
    if input.dim() != 4 or grid.dim() != 4 or grid.size(-1) != 2 or query.dim() != 3 or query.size(-1) != 2:
        raise RuntimeError("splat2d_lookup: expected input (N, C, H, W), grid (N, Hg, Wg, 2), query (N, P, 2)")
    n, c, h, w = input.shape
    if grid.size(0) != n or query.size(0) != n or values.shape[:2] != query.shape[:2] or values.size(2) != c or sigma.size(0) != n:
        raise RuntimeError("splat2d_lookup: batch / point / channel counts disagree")
    if c > 3:
        raise RuntimeError("splat2d_lookup: C <= 3 (RGB colours or a 1-channel mask)")
    input, grid, query, values, sigma = [t.float().contiguous() for t in (input, grid, query, values, sigma)]
    lib = _lib.load()
    out = torch.empty_like(input)
    points = torch.empty_like(query)
    ws = torch.empty(max(1, lib.gg_splat2d_workspace(n, c, h, w) // 4), dtype=torch.float32, device=input.device)
    rc = lib.gg_splat2d_lookup_forward(out.data_ptr(), points.data_ptr(), ws.data_ptr(), input.data_ptr(), grid.data_ptr(),
                                       query.data_ptr(), values.data_ptr(), sigma.data_ptr(), n, query.size(1), c, h, w,
                                       grid.size(1), grid.size(2), (res - 1) / res, float(out_res - 1),
                                       1 if soft_normalize else 0, _lib.stream())
    _lib.check(rc, "gg_splat2d_lookup_forward")
    return out, points


def nn_argmin(grid, points):
    """index[n, p] = argmin_{hw} |points[n, p]|^2 + |grid[n, hw]|^2 - 2 grid[n, hw] . points[n, p]  (first minimum), the
    brute-force search of `congeal_points` (spatial_transformer.py:655-668) without the (N, H, W, P) distance tensor.
    grid: (N, H, W, 2) or (N, HW, 2); points: (N, P, 2).  -> (N, P) int64."""
    _lib.require_cuda(grid, points)
    n = grid.size(0)
    g = grid.reshape(n, -1, 2).float().contiguous()
    pts = points.float().contiguous()
    if pts.dim() != 3 or pts.size(0) != n or pts.size(2) != 2:
        raise RuntimeError("nn_argmin: points must be (N, P, 2)")
    lib = _lib.load()
    p = pts.size(1)
    index = torch.empty((n, p), dtype=torch.int64, device=g.device)
    ws = torch.empty(max(1, lib.gg_nn_argmin_workspace(n, p) // 8), dtype=torch.int64, device=g.device)
    _lib.check(lib.gg_nn_argmin(index.data_ptr(), ws.data_ptr(), g.data_ptr(), pts.data_ptr(), n, p, g.size(1), _lib.stream()),
               "gg_nn_argmin")
    return index
