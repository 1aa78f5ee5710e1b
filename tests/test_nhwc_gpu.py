"""GPU parity of the channels-last (NHWC) kernel family (csrc/nhwc.cu) against the oracle -- same tolerances as NCHW."""
import pytest
import torch

from conftest import assert_close
from oracle import stylegan2_ops as so

pytestmark = pytest.mark.gpu
DEV = "cuda"
CL = torch.channels_last


def _cl(t):
    return t.to(DEV).contiguous(memory_format=CL)


@pytest.mark.parametrize("shape,pad,kind", [((2, 32, 65, 65), (1, 1), "1331"), ((2, 64, 129, 129), (1, 1), "1331"),
                                            ((1, 32, 40, 50), (2, 2), "1331"), ((2, 32, 17, 17), (1, 1), "1331"),
                                            ((1, 64, 64, 64), (2, 2), "rand"), ((1, 32, 70, 9), (2, 1), "rand3")])
def test_upfirdn2d_nhwc(shape, pad, kind):
    from gangealing_b200 import op
    g = torch.Generator().manual_seed(shape[2])
    x = torch.randn(*shape, generator=g)
    k = so.make_kernel([1, 3, 3, 1]) * 4 if kind == "1331" else torch.randn(3 if kind == "rand3" else 4, 4 if kind == "rand" else 3, generator=g)
    xg = _cl(x).requires_grad_(True)
    y = op.upfirdn2d(xg, k.to(DEV), pad=pad)
    assert y.is_contiguous(memory_format=CL)
    xo = x.clone().requires_grad_(True)
    yo = so.upfirdn2d_ref(xo, k, pad=pad)
    assert_close(y, yo, rtol=1e-4, what="fwd")
    go = torch.randn(yo.shape, generator=g)
    (gxo,) = torch.autograd.grad(yo, xo, go)
    (gx,) = torch.autograd.grad(y, xg, _cl(go))
    assert_close(gx, gxo, rtol=1e-4, what="grad")


@pytest.mark.parametrize("shape,pad", [((2, 32, 65, 65), (1, 1)), ((2, 64, 129, 129), (1, 1)), ((3, 64, 9, 9), (1, 1)),
                                       ((1, 32, 40, 50), (2, 2))])
def test_blur_noise_bias_act_nhwc(shape, pad):
    from gangealing_b200 import op
    g = torch.Generator().manual_seed(22)
    n, c, h, w = shape
    k = so.make_kernel([1, 3, 3, 1]) * 4
    oh, ow = h + pad[0] + pad[1] - 3, w + pad[0] + pad[1] - 3
    x = torch.randn(*shape, generator=g)
    noise = torch.randn(n, 1, oh, ow, generator=g)
    nw = torch.randn(1, generator=g)
    b = torch.randn(c, generator=g)
    rs = torch.rand(n, c, generator=g) + 0.5
    go = torch.randn(n, c, oh, ow, generator=g)
    leaves_o = [t.clone().requires_grad_(True) for t in (x, noise, nw, b, rs)]
    yo = so.blur_noise_bias_act_ref(leaves_o[0], k, pad, leaves_o[1], leaves_o[2], leaves_o[3], row_scale=leaves_o[4])
    grads_o = torch.autograd.grad(yo, leaves_o, go)
    leaves = [_cl(x).requires_grad_(True)] + [t.to(DEV).requires_grad_(True) for t in (noise, nw, b, rs)]
    y = op.blur_noise_bias_act(leaves[0], k.to(DEV), pad, leaves[1], leaves[2], leaves[3], row_scale=leaves[4])
    assert y.is_contiguous(memory_format=CL)
    assert_close(y, yo, rtol=1e-4, what="fwd")
    grads = torch.autograd.grad(y, leaves, _cl(go))
    for a, e, nm in zip(grads, grads_o, ("x", "noise", "noise_weight", "bias", "row_scale")):
        assert_close(a, e, rtol=3e-4, what="grad " + nm)


@pytest.mark.parametrize("shape", [(2, 128, 257, 257), (1, 128, 257, 65), (32, 64, 33, 33)])
def test_blur_noise_bias_act_nhwc_equals_nchw_path_at_full_size(shape):
    """Full-size planes: the channels-last path against the (oracle-verified) NCHW kernels on the same device inputs.
    (An oracle comparison at 16M outputs trips over pre-activations within one ulp of 0, whose leaky-relu slope flips.)"""
    from gangealing_b200 import op
    g = torch.Generator().manual_seed(3)
    n, c, h, w = shape
    k = (so.make_kernel([1, 3, 3, 1]) * 4).to(DEV)
    x = torch.randn(*shape, generator=g).to(DEV)
    noise = torch.randn(n, 1, h - 1, w - 1, generator=g).to(DEV)
    nw, b = torch.randn(1, generator=g).to(DEV), torch.randn(c, generator=g).to(DEV)
    rs = (torch.rand(n, c, generator=g) + 0.5).to(DEV)
    go = torch.randn(n, c, h - 1, w - 1, generator=g).to(DEV)
    res = []
    for cl in (False, True):
        leaves = [(x.contiguous(memory_format=CL) if cl else x.clone()).requires_grad_(True)] + \
                 [t.clone().requires_grad_(True) for t in (noise, nw, b, rs)]
        y = op.blur_noise_bias_act(leaves[0], k, (1, 1), leaves[1], leaves[2], leaves[3], row_scale=leaves[4])
        assert y.is_contiguous(memory_format=CL) == cl or n * c == 0
        res.append([y] + list(torch.autograd.grad(y, leaves, go.contiguous(memory_format=CL) if cl else go)))
    for a, e, nm in zip(res[1], res[0], ("fwd", "x", "noise", "noise_weight", "bias", "row_scale")):
        assert_close(a, e, rtol=2e-5, what=nm)


@pytest.mark.parametrize("shape", [(2, 32, 16, 16), (3, 512, 4, 4), (2, 128, 64, 64), (1, 8, 33, 31)])
def test_elementwise_family_nhwc(shape):
    from gangealing_b200 import op
    from gangealing_b200.op.modconv import channel_scale
    g = torch.Generator().manual_seed(5)
    n, c, h, w = shape
    x = torch.randn(*shape, generator=g)
    noise = torch.randn(n, 1, h, w, generator=g)
    nw, b = torch.randn(1, generator=g), torch.randn(c, generator=g)
    s = torch.randn(n, c, generator=g)
    go = torch.randn(*shape, generator=g)
    # noise_bias_act with row scale
    lo = [t.clone().requires_grad_(True) for t in (x, b, s)]
    yo = so.noise_bias_act_ref(lo[0] * lo[2][:, :, None, None], noise, nw, lo[1])
    g_o = torch.autograd.grad(yo, lo, go)
    lg = [_cl(x).requires_grad_(True), b.to(DEV).requires_grad_(True), s.to(DEV).requires_grad_(True)]
    y = op.noise_bias_act(lg[0], noise.to(DEV), nw.to(DEV), lg[1], row_scale=lg[2])
    assert y.is_contiguous(memory_format=CL)
    assert_close(y, yo, rtol=1e-5, what="noise_bias_act")
    g_g = torch.autograd.grad(y, lg, _cl(go))
    for a, e, nm in zip(g_g, g_o, ("x", "bias", "row_scale")):
        assert_close(a, e, rtol=2e-4, what="noise_bias_act grad " + nm)
    # fused_leaky_relu (FusedLeakyReLU of the STN trunk) on channels-last input
    lo = [t.clone().requires_grad_(True) for t in (x, b)]
    yo = so.fused_leaky_relu_ref(lo[0], lo[1])
    g_o = torch.autograd.grad(yo, lo, go)
    lg = [_cl(x).requires_grad_(True), b.to(DEV).requires_grad_(True)]
    y = op.fused_leaky_relu(lg[0], lg[1])
    assert_close(y, yo, rtol=1e-5, what="fused_leaky_relu")
    g_g = torch.autograd.grad(y, lg, _cl(go))
    assert_close(g_g[0], g_o[0], rtol=1e-5, what="flr gx")
    assert_close(g_g[1], g_o[1], rtol=2e-4, what="flr gbias")
    # channel_scale
    lo = [t.clone().requires_grad_(True) for t in (x, s)]
    yo = lo[0] * lo[1][:, :, None, None]
    g_o = torch.autograd.grad(yo, lo, go)
    lg = [_cl(x).requires_grad_(True), s.to(DEV).requires_grad_(True)]
    y = channel_scale(lg[0], lg[1])
    assert_close(y, yo, rtol=1e-6, what="channel_scale")
    g_g = torch.autograd.grad(y, lg, _cl(go))
    assert_close(g_g[0], g_o[0], rtol=1e-6)
    assert_close(g_g[1], g_o[1], rtol=2e-4)


def test_generator_channels_last_matches_nchw_and_fixture():
    from conftest import load_golden
    from gangealing_b200.stylegan2 import Generator
    from oracle import opset
    blob = load_golden("networks")
    g = opset.fill_parameters(Generator(32, 32, 2, channel_multiplier=2).eval(), 1).to(DEV)
    noise = [blob["gen.noise%d" % i].to(DEV) for i in range(g.num_layers)]
    z = blob["gen.z"].to(DEV)
    with torch.no_grad():
        a, _ = g([z], noise=noise)
        g.channels_last = True
        b, _ = g([z], noise=noise)
    assert_close(b, a, rtol=1e-3, what="NHWC vs NCHW generator")
    assert_close(b, blob["gen.image"], rtol=1e-3, what="NHWC generator vs reference fixture")


@pytest.mark.parametrize("transform", ["similarity", "flow"])
def test_stn_trunk_channels_last_matches_nchw(transform):
    from gangealing_b200.stn import SpatialTransformer
    from oracle import opset
    torch.backends.cudnn.allow_tf32 = False
    try:
        stn = opset.fill_parameters(SpatialTransformer(64, 64, channel_multiplier=0.5, transform=transform), 3).to(DEV)
        img = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1)).to(DEV)
        res = []
        for cl in (False, True):
            stn.channels_last = cl
            stn.zero_grad()
            out, grid = stn(img, return_flow=True)
            (out.square().mean() + grid.square().mean()).backward()
            res.append((out.detach(), grid.detach(), [p.grad.clone() for p in stn.parameters() if p.grad is not None]))
        assert_close(res[1][0], res[0][0], rtol=5e-4, what="warped image")   # cuDNN picks other algorithms per layout
        assert_close(res[1][1], res[0][1], rtol=5e-4, what="grid")
        assert len(res[0][2]) == len(res[1][2]) > 0
        # all parameter gradients as one vector: tiny-magnitude tensors are judged on the scale of the whole gradient
        assert_close(torch.cat([a.flatten() for a in res[1][2]]), torch.cat([e.flatten() for e in res[0][2]]), rtol=2e-3,
                     what="parameter grads")
    finally:
        torch.backends.cudnn.allow_tf32 = True


@pytest.mark.parametrize("shape,with_skip", [((2, 32, 8, 8), True), ((3, 128, 33, 31), True), ((2, 512, 4, 4), False),
                                             ((2, 64, 64, 64), True), ((1, 256, 16, 16), False)])
def test_to_rgb_nhwc_matches_the_reference_formulation(shape, with_skip):
    """Fused to-RGB (1x1 modulated conv, no demod, + bias + skip) on channels-last input vs the oracle's grouped
    formulation of reference networks.py:389-405; gradients w.r.t. activation, style, bias and skip."""
    from gangealing_b200.op.modconv import modulated_conv2d
    from oracle import opset
    g = torch.Generator().manual_seed(shape[1])
    n, c, h, w = shape
    x = torch.randn(*shape, generator=g)
    weight = torch.randn(1, 3, c, 1, 1, generator=g)
    style = torch.randn(n, c, generator=g)
    bias = torch.randn(1, 3, 1, 1, generator=g)
    skip = torch.randn(n, 3, h, w, generator=g) if with_skip else None
    go = torch.randn(n, 3, h, w, generator=g)
    scale = 1.0 / c ** 0.5
    lo = [t.clone().requires_grad_(True) for t in (x, style, bias)] + ([skip.clone().requires_grad_(True)] if with_skip else [])
    yo, _ = opset.cpu_ops().modulated_conv2d(lo[0], weight, lo[1], scale, False, False, 0, 1e-8, bias=lo[2],
                                             skip=lo[3] if with_skip else None)
    g_o = torch.autograd.grad(yo, lo, go)
    lg = [_cl(x).requires_grad_(True)] + [t.to(DEV).requires_grad_(True) for t in (style, bias)] + \
         ([skip.to(DEV).requires_grad_(True)] if with_skip else [])
    y, d = modulated_conv2d(lg[0], weight.to(DEV), lg[1], scale, False, False, 0, 1e-8, bias=lg[2],
                            skip=lg[3] if with_skip else None)
    assert d is None
    assert_close(y, yo, rtol=2e-5, what="to_rgb fwd")
    g_g = torch.autograd.grad(y, lg, go.to(DEV))
    assert g_g[0].is_contiguous(memory_format=CL)
    for a, e, nm in zip(g_g, g_o, ("x", "style", "bias", "skip")):
        assert_close(a, e, rtol=2e-4, what="to_rgb grad " + nm)
