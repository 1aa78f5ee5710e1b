"""CPU: the oracle restatement reproduces the fixtures generated from the reference (oracle/make_golden.py)."""
import torch

from conftest import assert_close, golden_cases, load_golden
from oracle import stylegan2_ops as so


def test_upfirdn2d_oracle_matches_reference_fixtures():
    blob = load_golden("upfirdn2d")
    names = golden_cases(blob)
    assert len(names) >= 12
    for name in names:
        up, down, p0, p1 = [int(v) for v in blob[name + ".cfg"]]
        y = so.upfirdn2d_ref(blob[name + ".x"], blob[name + ".k"], up=up, down=down, pad=(p0, p1))
        assert_close(y, blob[name + ".y"], rtol=1e-6, what=name)


def test_fused_act_oracle_matches_reference_fixtures():
    blob = load_golden("fused_act")
    for name in golden_cases(blob):
        x, b, g = blob[name + ".x"], blob[name + ".b"], blob[name + ".g"]
        y = so.fused_leaky_relu_ref(x, b)
        assert_close(y, blob[name + ".y"], rtol=1e-6, what=name + " fwd")
        gx, gb = so.fused_leaky_relu_backward_ref(g, y)
        assert_close(gx, blob[name + ".gx"], rtol=1e-6, what=name + " gx")
        assert_close(gb, blob[name + ".gb"], rtol=1e-5, what=name + " gb")


def test_fused_bias_act_table():
    # act/grad table of fused_bias_act_kernel.cu:28-47
    x = torch.tensor([[-2.0, 3.0], [0.5, -1.0]])
    ref = torch.tensor([[1.0, -1.0], [-1.0, 1.0]])
    b = torch.tensor([1.0, -1.0])
    assert torch.equal(so.fused_bias_act_ref(x, None, None, 1, 0, 0.2, 2.0), x * 2)
    assert torch.equal(so.fused_bias_act_ref(x, b, None, 1, 1, 0.2, 1.0), x + b)
    y = so.fused_bias_act_ref(x, None, None, 3, 0, 0.5, 1.0)
    assert torch.equal(y, torch.tensor([[-1.0, 3.0], [0.5, -0.5]]))
    y = so.fused_bias_act_ref(x, None, ref, 3, 1, 0.5, 1.0)
    assert torch.equal(y, torch.tensor([[-2.0, 1.5], [0.25, -1.0]]))
    assert torch.equal(so.fused_bias_act_ref(x, None, ref, 3, 2, 0.5, 1.0), torch.zeros(2, 2))


def test_upfirdn2d_is_linear_and_shift_consistent():
    # size-independent properties also used on the GPU at full size
    g = torch.Generator().manual_seed(0)
    k = so.make_kernel([1, 3, 3, 1])
    a, b = torch.randn(1, 2, 20, 20, generator=g), torch.randn(1, 2, 20, 20, generator=g)
    ya = so.upfirdn2d_ref(a, k, pad=(2, 1))
    yb = so.upfirdn2d_ref(b, k, pad=(2, 1))
    assert_close(so.upfirdn2d_ref(2 * a - 3 * b, k, pad=(2, 1)), 2 * ya - 3 * yb, rtol=1e-5)
    ones = torch.ones(1, 1, 16, 16)
    y = so.upfirdn2d_ref(ones, k, pad=(2, 1))
    assert_close(y[:, :, 3:-3, 3:-3], torch.ones(1, 1, 10, 10), rtol=1e-6)  # unit DC gain in the interior
