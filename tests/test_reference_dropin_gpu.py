"""GPU: the UNMODIFIED reference networks executing on libgg_b200's kernels through `gangealing_b200.compat`
("models/stylegan2 and models/spatial_transformers drop in unchanged above the kernel boundary").

The reference checkout does not exist on the GPU box, and no reference source may enter this repo; what travels is
oracle/_ref/refpy -- the reference's own `models/` package byte-compiled by oracle/build_ref.py (a compiled output,
like the reference CUDA kernels in oracle/_ref/*.so).  The modules that ARE the kernel boundary
(models/stylegan2/op, models/spatial_transformers/antialiased_sampling.py) are not in that tree: the shim supplies
them.  Expected values: tests/golden/networks.npz, written by the reference on its CPU path (oracle/make_golden.py).
"""
import sys

import pytest
import torch

from conftest import assert_close, load_golden
from oracle import build_ref, opset

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.refpy_available(), reason="oracle/_ref/refpy not built (python -m oracle.build_ref)")]
DEV = "cuda"


@pytest.fixture()
def reference():
    """The reference's `models` package, imported from the byte-compiled tree above the compat shim."""
    import gangealing_b200.compat as compat
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")}
    for k in saved:
        del sys.modules[k]
    compat.install(force=True)
    sys.path.insert(0, build_ref.REFPY)
    try:
        import models
        assert models.__file__.startswith(build_ref.REFPY), models.__file__
        yield models
    finally:
        sys.path.remove(build_ref.REFPY)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_unmodified_reference_generator_on_gpu_matches_reference_fixture(reference):
    import gangealing_b200.op as op
    from models.stylegan2 import networks
    assert networks.upfirdn2d is op.upfirdn2d and networks.FusedLeakyReLU is op.FusedLeakyReLU  # our kernels underneath
    blob = load_golden("networks")
    g = opset.fill_parameters(reference.Generator(32, 32, 2, channel_multiplier=2).eval(), 1).to(DEV)
    noise = [blob["gen.noise%d" % i].to(DEV) for i in range(g.num_layers)]
    with torch.no_grad():
        img, lat = g([blob["gen.z"].to(DEV)], noise=noise, return_latents=True)
    assert_close(lat, blob["gen.latent"], rtol=1e-4, what="latent")
    assert_close(img, blob["gen.image"], rtol=1e-3, what="image")


@pytest.mark.parametrize("transforms", [("similarity",), ("similarity", "flow")])
def test_unmodified_reference_stn_on_gpu_matches_reference_fixture(reference, transforms):
    from gangealing_b200.stn import sampling
    blob = load_golden("networks")
    tag = "stn_" + "_".join(transforms)
    stn = reference.get_stn(list(transforms), flow_size=64, supersize=128, channel_multiplier=0.5, num_heads=1).eval()
    first = stn if len(transforms) == 1 else stn.stns[0]
    assert isinstance(first.warp_head.warper, sampling.MipmapWarp)      # the fused sampler, not the reference's
    opset.fill_parameters(stn, 3, gain=0.3).to(DEV)
    with torch.no_grad():
        out, grid, fm = stn(blob[tag + ".x"].to(DEV), return_warp=True, return_flow=True, padding_mode="reflection")
    assert_close(grid, blob[tag + ".grid"], rtol=1e-3, what="grid")
    assert_close(fm, blob[tag + ".fm"], rtol=1e-3, what="flow/matrix")
    assert_close(out, blob[tag + ".out"], rtol=2e-3, what="warped image")


def test_unmodified_reference_loss_and_gradients_equal_this_repos_mirror(reference):
    """The reference's `gangealing_loss` (its Generator, ComposedSTN, DirectionInterpolator, BilinearDownsample callers)
    through the shim vs this repo's host-side mirror: same weights, same CUDA RNG stream -> same loss and gradients."""
    from gangealing_b200.stn import BilinearDownsample, get_stn
    from gangealing_b200.stylegan2 import Generator
    from gangealing_b200.training.latent_learner import DirectionInterpolator
    from gangealing_b200.training.losses import gangealing_loss

    def mse(a, b):
        return (a - b).pow(2).mean(dim=(1, 2, 3))

    def build(G, STN, LL, DOWN):
        g = opset.fill_parameters(G(128, 512, 2, channel_multiplier=1).eval(), 11).to(DEV)
        for prm in g.parameters():
            prm.requires_grad = False
        stn = STN(["similarity", "flow"], flow_size=64, supersize=128, channel_multiplier=0.25, num_heads=1)
        opset.fill_parameters(stn, 12, gain=0.2).to(DEV)
        ll = opset.fill_parameters(LL(None, 2, 3, g.n_latent, num_heads=1), 13, gain=0.5).to(DEV)
        return g, stn, ll, DOWN(2, 3).to(DEV)

    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        res = []
        for loss_fn, parts in ((reference.gangealing_loss, build(reference.Generator, reference.get_stn,
                                                                 reference.DirectionInterpolator, reference.BilinearDownsample)),
                               (gangealing_loss, build(Generator, get_stn, DirectionInterpolator, BilinearDownsample))):
            g, stn, ll, down = parts
            torch.manual_seed(77)
            loss, delta = loss_fn(g, stn, ll, mse, down, 0.6, 2, 512, False, DEV, sample_from_full_res=True,
                                  padding_mode="reflection")
            grads = torch.autograd.grad(loss + 10.0 * reference.total_variation_loss(delta),
                                        list(stn.parameters()) + [ll.coefficients], allow_unused=True)
            res.append((loss.detach(), delta.detach(), [n for n, _ in stn.named_parameters()] + ["ll"], grads))
        (l_ref, d_ref, names, g_ref), (l_our, d_our, names2, g_our) = res
        assert names == names2                                      # identical module trees
        assert_close(l_our, l_ref, rtol=1e-4, what="loss")
        assert_close(d_our, d_ref, rtol=1e-4, what="delta flow")
        pairs = [(a, b) for a, b in zip(g_ref, g_our) if a is not None and b is not None]
        assert len(pairs) > 20 and all((a is None) == (b is None) for a, b in zip(g_ref, g_our))
        assert_close(torch.cat([b.flatten() for _, b in pairs]), torch.cat([a.flatten() for a, _ in pairs]), rtol=1e-3,
                     what="whole gradient")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
