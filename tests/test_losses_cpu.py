"""CPU: the callers of the hot path (pair sampling, gangealing_loss, gangealing_cluster_loss -- BASELINE configs 2 and 5,
shrunk) reproduce the reference end to end: same seeded weights, same global-RNG consumption order."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import opset

CPU = opset.cpu_ops()


def _mse(a, b):
    return (a - b).pow(2).mean(dim=(1, 2, 3))


@pytest.mark.parametrize("tag,heads,flips,full_res", [("uni", 1, False, False), ("cluster", 2, True, True)])
def test_losses_match_reference_fixture(tag, heads, flips, full_res):
    from gangealing_b200.stn import BilinearDownsample, get_stn
    from gangealing_b200.stylegan2 import Generator
    from gangealing_b200.training import (DirectionInterpolator, gangealing_cluster_loss, gangealing_loss,
                                          total_variation_loss)
    blob = load_golden("losses")
    gen_size = 128 if full_res else 64
    g = opset.fill_parameters(Generator(gen_size, 512, 2, channel_multiplier=1, ops=CPU).eval(), 11)
    for prm in g.parameters():
        prm.requires_grad = False
    stn = get_stn(["similarity", "flow"], flow_size=64, supersize=gen_size, channel_multiplier=0.25, num_heads=heads, ops=CPU)
    opset.fill_parameters(stn, 12, gain=0.2)
    ll = DirectionInterpolator(None, 2, 3, g.n_latent, num_heads=heads)
    opset.fill_parameters(ll, 13, gain=0.5)
    resize = BilinearDownsample(2, 3, ops=CPU) if full_res else torch.nn.Sequential()
    torch.manual_seed(1234)
    if heads == 1:
        loss, delta = gangealing_loss(g, stn, ll, _mse, resize, 0.6, 2, 512, False, "cpu", sample_from_full_res=full_res,
                                      padding_mode="reflection")
    else:
        loss, delta = gangealing_cluster_loss(g, stn, ll, _mse, resize, 0.6, 2, 512, False, heads, flips, "cpu",
                                              sample_from_full_res=full_res, padding_mode="reflection")
    tv = total_variation_loss(delta)
    assert_close(loss, blob[tag + ".loss"], rtol=1e-4, what="loss")
    assert_close(tv, blob[tag + ".tv"], rtol=1e-4, what="tv")
    assert_close(delta, blob[tag + ".delta"], rtol=1e-4, what="delta_flow")
    params = dict(stn.named_parameters())
    params["ll.coefficients"] = ll.coefficients
    wanted = [k[len(tag) + 6:] for k in blob if k.startswith(tag + ".grad.")]
    assert len(wanted) >= 3
    grads = torch.autograd.grad(loss + 10.0 * tv, [params[n] for n in wanted])
    for n, gr in zip(wanted, grads):
        assert_close(gr, blob[tag + ".grad." + n], rtol=2e-3, what="grad " + n)
