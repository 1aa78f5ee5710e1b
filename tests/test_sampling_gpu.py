"""GPU parity of the fused antialiased sampler (csrc/warp.cu) against the reference fixtures and the oracle."""
import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close, golden_cases, load_golden
from oracle import sampling as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _stn():
    from gangealing_b200 import stn
    return stn


def test_mipmap_warp_golden_forward_backward_and_levels():
    stn = _stn()
    blob = load_golden("mipmap_warp")
    mw = stn.MipmapWarp(3.5).to(DEV)
    for name in golden_cases(blob):
        mode = S.PAD_MODES[int(blob[name + ".mode"])]
        x = blob[name + ".x"].to(DEV).requires_grad_(True)
        grid = blob[name + ".grid"].to(DEV).requires_grad_(True)
        y = mw(x, grid, padding_mode=mode)
        assert_close(y, blob[name + ".y"], rtol=1e-4, what=name + " fwd")
        gx, gg = torch.autograd.grad(y, [x, grid], blob[name + ".go"].to(DEV))
        assert_close(gx, blob[name + ".gx"], rtol=1e-4, what=name + " gx")
        assert_close(gg, blob[name + ".ggrid"], rtol=1e-3, what=name + " ggrid")
        lv = (mw.levels_map * 2.5).cpu()
        ref_lv = blob[name + ".levels"]
        assert_close(lv, ref_lv, atol=2e-6, what=name + " levels")
        # integer level indices: exact wherever the level is not within float noise of an integer
        safe = (ref_lv - ref_lv.round()).abs() > 1e-5
        assert torch.equal(lv.floor()[safe], ref_lv.floor()[safe]) and torch.equal(lv.ceil()[safe], ref_lv.ceil()[safe])
        # plain Warp
        w = stn.Warp()
        x2 = blob[name + ".x"].to(DEV).requires_grad_(True)
        g2 = blob[name + ".grid"].to(DEV).requires_grad_(True)
        yw = w(x2, g2, padding_mode=mode)
        assert_close(yw, blob[name + ".warp_y"], rtol=1e-5, what=name + " warp fwd")
        gxw, ggw = torch.autograd.grad(yw, [x2, g2], blob[name + ".go"].to(DEV))
        assert_close(gxw, blob[name + ".warp_gx"], rtol=1e-4, what=name + " warp gx")
        assert_close(ggw, blob[name + ".warp_ggrid"], rtol=1e-3, what=name + " warp ggrid")


@pytest.mark.parametrize("size,res,mode", [(128, 128, "border"), (256, 128, "reflection"), (450, 128, "border"),
                                           (512, 512, "border"), (64, 96, "zeros")])
def test_mipmap_warp_vs_oracle_training_shapes(size, res, mode):
    stn = _stn()
    g = torch.Generator().manual_seed(size + res)
    n = 3
    x = torch.randn(n, 3, size, size, generator=g)
    theta = torch.tensor([[1.0, 0.0, 0.0, 0.0, 1.0, 0.0], [1.9, 0.6, 0.1, -0.6, 1.9, -0.1], [3.0, 0.0, 0.2, 0.0, 3.0, 0.0]]).reshape(3, 2, 3)
    coarse = torch.randn(n, 2, 6, 6, generator=g)
    grid = F.affine_grid(theta, (n, 3, res, res), align_corners=False) + \
        0.08 * F.interpolate(coarse, size=(res, res), mode="bicubic", align_corners=False).permute(0, 2, 3, 1)
    yo, aux = S.mipmap_warp_ref(x, grid, 3.5, 0.0, mode, return_aux=True)
    mw = stn.MipmapWarp(3.5).to(DEV)
    y = mw(x.to(DEV), grid.to(DEV), padding_mode=mode)
    assert_close(y, yo, rtol=1e-4, what="fwd")
    assert_close(mw.levels_map.cpu() * 2.5, aux["levels"], atol=5e-6, what="levels")


def test_mipmap_warp_min_level_and_warp_match_torch_grid_sample():
    stn = _stn()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 32, 40, generator=g)  # non-square, 5 channels: plain Warp only
    grid = torch.rand(2, 17, 23, 2, generator=g) * 2.6 - 1.3
    for mode in S.PAD_MODES:
        y = stn.Warp()(x.to(DEV), grid.to(DEV), padding_mode=mode)
        assert_close(y, F.grid_sample(x, grid, padding_mode=mode, align_corners=False), rtol=1e-5, what=mode)
    x = torch.randn(2, 3, 64, 64, generator=g)
    grid = F.affine_grid(torch.eye(2, 3)[None].repeat(2, 1, 1) * 1.7, (2, 3, 32, 32), align_corners=False)
    mw = stn.MipmapWarp(3.5).to(DEV)
    for min_level in (0.5, 2.0):
        y = mw(x.to(DEV), grid.to(DEV), min_level=min_level)
        assert_close(y, S.mipmap_warp_ref(x, grid, 3.5, min_level, "border"), rtol=1e-4, what="min_level %g" % min_level)


def test_mipmap_warp_low_precision_and_errors():
    stn = _stn()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 64, 64, generator=g)
    grid = F.affine_grid(torch.tensor([[[2.0, 0.1, 0.0], [-0.1, 2.0, 0.0]]]).repeat(2, 1, 1), (2, 3, 32, 32), align_corners=False)
    yo = S.mipmap_warp_ref(x.bfloat16().float(), grid, 3.5, 0.0, "border")
    y = stn.MipmapWarp(3.5).to(DEV)(x.bfloat16().to(DEV), grid.to(DEV))
    assert y.dtype == torch.bfloat16
    assert_close(y, yo, rtol=1.6e-2)
    with pytest.raises(RuntimeError):
        stn.MipmapWarp(3.5)(x, grid)  # CPU tensors
    with pytest.raises(RuntimeError):
        stn.Warp()(x.to(DEV), grid.to(DEV), padding_mode="wrap")
    with pytest.raises(RuntimeError):
        stn.Warp()(x.to(DEV), grid[:1].to(DEV))


def test_bilinear_downsample_golden():
    stn = _stn()
    blob = load_golden("bilinear_downsample")
    for stride in (2, 4):
        y = stn.BilinearDownsample(stride, 3).to(DEV)(blob["x"].to(DEV))
        assert_close(y, blob["s%d.y" % stride], rtol=1e-5)


@pytest.mark.parametrize("shape,stride", [((2, 3, 32, 32), 2), ((1, 3, 33, 29), 2), ((2, 3, 64, 48), 4), ((1, 2, 40, 40), 8),
                                          ((2, 3, 9, 7), 1), ((1, 3, 21, 30), 3), ((4, 3, 256, 256), 2)])
def test_bilinear_downsample_forward_and_adjoint(shape, stride):
    """One-kernel BilinearDownsample vs the oracle (reference antialiased_sampling.py:241-256): forward, and the
    gather-form backward against autograd of the reference formulation."""
    stn = _stn()
    g = torch.Generator().manual_seed(shape[2] * 7 + stride)
    x = torch.randn(*shape, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = S.bilinear_downsample_ref(xo, stride)
    go = torch.randn(yo.shape, generator=g)
    (gxo,) = torch.autograd.grad(yo, xo, go)
    mod = stn.BilinearDownsample(stride, shape[1]).to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = mod(xg)
    assert y.shape == yo.shape
    assert_close(y, yo, rtol=1e-5, what="forward")
    (gx,) = torch.autograd.grad(y, xg, go.to(DEV))
    assert_close(gx, gxo, rtol=1e-5, what="backward")
    lhs = (y.detach().double() * go.to(DEV).double()).sum()
    rhs = (xg.detach().double() * gx.double()).sum()
    assert abs(lhs - rhs) <= 1e-6 * (y.detach().double() * go.to(DEV).double()).abs().sum()


def test_bilinear_downsample_errors():
    stn = _stn()
    with pytest.raises(RuntimeError):
        stn.BilinearDownsample(2, 3)(torch.zeros(1, 3, 8, 8))             # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        stn.BilinearDownsample(4, 3).to(DEV)(torch.zeros(1, 3, 2, 2, device=DEV))   # plane not larger than stride/2
    y = stn.BilinearDownsample(2, 3).to(DEV)(torch.zeros(0, 3, 8, 8, device=DEV))
    assert y.shape == (0, 3, 4, 4)


@pytest.mark.parametrize("mode", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("hs,ws", [(128, 128), (64, 256)])
def test_sampler_integer_work_is_bit_exact(mode, hs, ws):
    """The sampler's INTEGER work compared as integers (north_star: "bit-exact for sampling-grid integer/index work"):
    bilinear corner indices (x0, y0) after the padding-mode transform and the floor / ceil level indices, exported by
    gg_warp_sample_indices from the device functions the sampling kernels call, vs the oracle's
    (oracle/sampling.py grid_sample_bilinear / mipmap_levels = ATen grid_sampler + antialiased_sampling.py:197-229).

    (a) DYADIC grid: every coordinate k/512 -- all of ((g+1)*size-1)/2, the reflection fold and the floor are exact in
        fp32 on both sides (no rounding, so FMA contraction cannot matter): ALL pixels must agree, including the exact
        ties where the source coordinate IS an integer and the far out-of-range coordinates of every padding mode.
    (b) random smooth grid: exact wherever the oracle's coordinate is not within 1e-4 px of an integer (there the last
        ulp of the fp32 evaluation order decides, on the GPU as in ATen's own CUDA kernel); that exempt set must be tiny."""
    from gangealing_b200.stn import sampling as GS
    g = torch.Generator().manual_seed(hs + ws)
    n, ho, wo = 2, 96, 80
    k = torch.randint(-900, 900, (n, ho, wo, 2), generator=g)
    k[0, :8, :8] = torch.tensor([4, -4])                 # exact ties: coordinate lands on an integer pixel
    k[0, 8:16, :8] = torch.tensor([-512, 512])           # the two borders of the normalised range
    dyadic = k.float() / 512.0
    theta = torch.tensor([[[0.9, 0.2, 0.05], [-0.15, 1.3, -0.1]], [[2.2, 0.0, 0.3], [0.1, 1.9, 0.0]]])
    smooth = F.affine_grid(theta, (n, 1, ho, wo), align_corners=False) + 0.03 * torch.randn(n, ho, wo, 2, generator=g)
    img = torch.zeros(n, 1, hs, ws)
    for name, grid in (("dyadic", dyadic), ("smooth", smooth)):
        got = GS.sample_indices(grid.to(DEV), (hs, ws), 3.5, 0.0, mode).cpu()
        _, (x0, y0) = S.grid_sample_bilinear(img, grid, mode)
        lv = S.mipmap_levels(grid, hs, ws, 3.5)
        ix = S.source_index(grid[..., 0], ws, mode)
        iy = S.source_index(grid[..., 1], hs, mode)
        if name == "dyadic":
            corner_ok = torch.ones_like(x0, dtype=torch.bool)
        else:
            def decided(c, size):      # not within 1e-4 px of an integer -- or pinned to a border pixel by the clamp of the
                ok = (c - c.round()).abs() > 1e-4                      # border / reflection modes (an exact constant on both sides)
                if mode != "zeros":
                    ok |= (c == 0) | (c == size - 1)
                return ok
            corner_ok = decided(ix, ws) & decided(iy, hs)
            assert corner_ok.float().mean() > 0.995
        assert torch.equal(got[..., 0][corner_ok].long(), x0[corner_ok]), name + " x0"
        assert torch.equal(got[..., 1][corner_ok].long(), y0[corner_ok]), name + " y0"
        level_ok = (lv - lv.round()).abs() > 1e-5
        level_ok |= lv == 0                                 # the clamp: distance <= 1 px is level 0 exactly on both sides
        assert level_ok.float().mean() > 0.995
        assert torch.equal(got[..., 2][level_ok].long(), lv.floor().long()[level_ok]), name + " floor(level)"
        assert torch.equal(got[..., 3][level_ok].long(), lv.ceil().long()[level_ok]), name + " ceil(level)"
