"""Perceptual-loss front end: oracle vs the reference-generated fixture (CPU), the fused
channels-last kernels vs the oracle through the C ABI (GPU)."""
import pytest
import torch

from conftest import assert_close, golden_cases, load_golden
from oracle.perceptual import feature_distance_ref

DEV = "cuda"


def _case(blob, name):
    w = blob.get(name + ".weight")
    return blob[name + ".f0"], blob[name + ".f1"], w, blob[name + ".out"], blob[name + ".gout"], blob[name + ".g0"], blob[name + ".g1"]


def test_oracle_matches_reference_fixture():
    blob = load_golden("perceptual")
    names = golden_cases(blob)
    assert len(names) >= 5
    for name in names:
        f0, f1, w, out, gout, g0, g1 = _case(blob, name)
        a, b = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
        res = feature_distance_ref(a, b, w)
        assert res.shape == out.shape
        assert_close(res, out, rtol=1e-6, what=name + " out")
        ga, gb = torch.autograd.grad(res, [a, b], gout)
        assert_close(torch.nan_to_num(ga, nan=0.0), g0, rtol=1e-6, what=name + " g0")
        assert_close(gb, g1, rtol=1e-6, what=name + " g1")


def test_device_side_composite_is_the_same_formula_and_the_op_refuses_cpu_tensors():
    from gangealing_b200.op.feature_distance import _composite, feature_distance
    g = torch.Generator().manual_seed(2)
    a, b = torch.rand(2, 12, 5, 5, generator=g), torch.rand(2, 12, 5, 5, generator=g)
    w = torch.rand(12, generator=g)
    assert_close(_composite(a, b, w, 1e-10), feature_distance_ref(a, b, w), rtol=1e-6)   # the unsupported-layout route
    with pytest.raises(RuntimeError):
        feature_distance(a, b, w)                                                         # no CPU path in the product


@pytest.mark.gpu
def test_fused_kernels_match_reference_fixture():
    from gangealing_b200.op.feature_distance import feature_distance, _supported
    blob = load_golden("perceptual")
    for name in golden_cases(blob):
        f0, f1, w, out, gout, g0, g1 = _case(blob, name)
        a = f0.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        b = f1.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        if a.shape[2] * a.shape[3] > 1 and a.shape[1] > 1:
            assert _supported(a, b), name
        res = feature_distance(a, b, None if w is None else w.to(DEV))
        assert_close(res, out, rtol=1e-5, what=name + " out")
        ga, gb = torch.autograd.grad(res, [a, b], gout.to(DEV))
        assert ga.is_contiguous(memory_format=torch.channels_last)
        assert_close(ga, g0, rtol=1e-4, what=name + " g0")
        assert_close(gb, g1, rtol=1e-4, what=name + " g1")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 64, 64, 64), (2, 128, 33, 31), (3, 256, 16, 16), (2, 512, 8, 8), (1, 512, 1, 1),
                                   (2, 16, 9, 9)])
def test_fused_kernels_match_oracle(shape):
    from gangealing_b200.op.feature_distance import feature_distance
    g = torch.Generator().manual_seed(shape[1] + shape[2])
    f0 = torch.relu(torch.randn(*shape, generator=g))
    f1 = torch.relu(torch.randn(*shape, generator=g) + 0.2)
    go = torch.randn(shape[0], 1, 1, 1, generator=g)
    a, b = f0.clone().requires_grad_(True), f1.clone().requires_grad_(True)
    ro = feature_distance_ref(a, b)
    gao, gbo = torch.autograd.grad(ro, [a, b], go)
    x = f0.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = f1.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = feature_distance(x, y)
    assert_close(r, ro, rtol=1e-5, what="out")
    gx, gy = torch.autograd.grad(r, [x, y], go.to(DEV))
    assert_close(gx, gao, rtol=1e-4, what="g0")
    assert_close(gy, gbo, rtol=1e-4, what="g1")
    # symmetry and identity: d(a, b) == d(b, a), d(a, a) == 0 -- size-independent properties
    assert_close(feature_distance(y, x), r, rtol=1e-6)
    assert float(feature_distance(x, x).abs().max()) < 1e-12   # a*ia - b*ib contracts to an fma: one rounding residual


def test_whole_perceptual_loss_matches_the_reference_lpips_fixture():
    """PerceptualLoss (this repo's mirror of LPIPS(net='vgg', lpips=False, pnet_rand=True)/18, lpips.py:13-17) against
    the reference class run with the same seeded VGG16 weights: scaling layer, slice boundaries, distance, gradients."""
    from gangealing_b200.training.perceptual import PerceptualLoss
    from oracle import opset
    blob = load_golden("perceptual_loss")
    loss = opset.fill_convs_in_order(PerceptualLoss(ops=opset.cpu_ops()), 4242)
    in0 = blob["in0"].clone().requires_grad_(True)
    in1 = blob["in1"].clone().requires_grad_(True)
    val = loss(in0, in1)
    assert val.shape == blob["val"].shape
    assert_close(val, blob["val"], rtol=1e-5, what="perceptual distance")
    g0, g1 = torch.autograd.grad(val.sum(), [in0, in1])
    assert_close(g0, blob["g0"], rtol=1e-4, what="d/d in0")
    assert_close(g1, blob["g1"], rtol=1e-4, what="d/d in1")
