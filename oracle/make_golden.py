"""Generate tests/golden/*.npz from the reference itself and pin the oracle against it.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
For every case the script (1) runs the REFERENCE's own CPU implementation on seeded inputs, (2) asserts
that the oracle restatement reproduces it (tight tolerance / exact where integer), (3) stores inputs and
reference outputs as a small fixture.  Fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import refimport, stylegan2_ops as so  # noqa: E402


def _save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote %-40s %7.1f KB" % (os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


def _close(a, b, tol, what):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err <= tol * max(1.0, ref), "%s: oracle deviates from the reference by %g" % (what, err)


# ---------------------------------------------------------------------------------------------------
UPFIRDN_CASES = [
    # name, shape, kernel spec, up, down, pad
    ("g_blur_9", (2, 3, 9, 9), "1331x4", 1, 1, (1, 1)),         # Generator blur after up-conv (networks.py:199-205)
    ("g_blur_65", (1, 2, 65, 65), "1331x4", 1, 1, (1, 1)),
    ("stn_blur_pad22", (1, 2, 32, 32), "1331", 1, 1, (2, 2)),   # ResBlock conv2 blur (networks.py:606-611)
    ("stn_blur_pad11", (1, 2, 32, 32), "1331", 1, 1, (1, 1)),   # ResBlock skip blur
    ("stn_blur_128", (1, 1, 128, 128), "1331", 1, 1, (1, 1)),   # 127-wide output: ragged strips
    ("blur_bwd_pad22", (1, 2, 64, 64), "1331x4", 1, 1, (2, 2)),
    ("rgb_up2", (2, 3, 8, 8), "1331x4", 2, 1, (2, 1)),          # Upsample (networks.py:28-46)
    ("rgb_up2_bwd_dn2", (2, 3, 16, 16), "1331x4", 1, 2, (1, 1)),
    ("down2", (1, 2, 16, 16), "1331", 1, 2, (1, 1)),            # Downsample (networks.py:49-67)
    ("k5_up2_dn3", (1, 2, 11, 13), "rand5", 2, 3, (3, 2)),      # generic path
    ("k3_asym", (1, 2, 12, 10), "rand3", 1, 1, (1, 1)),         # flip matters
    ("k2", (1, 1, 7, 9), "rand2", 1, 1, (0, 1)),
    ("neg_pad", (1, 2, 20, 20), "1331", 1, 1, (-1, 2)),         # negative pad crops
    ("k43_rect", (1, 1, 15, 15), "rand43", 1, 1, (2, 1)),
    ("tall_bands", (1, 1, 300, 40), "1331", 1, 1, (2, 2)),      # several bands per plane
]


def _kernel(spec, gen):
    if spec.startswith("1331"):
        k = so.make_kernel([1, 3, 3, 1])
        return k * 4 if spec.endswith("x4") else k
    if spec == "rand43":
        return torch.randn(4, 3, generator=gen)
    n = int(spec[4:])
    return torch.randn(n, n, generator=gen)


def gen_upfirdn2d(ref_models):
    from models.stylegan2.op.upfirdn2d import upfirdn2d as ref_upfirdn2d
    out = {}
    for i, (name, shape, spec, up, down, pad) in enumerate(UPFIRDN_CASES):
        gen = torch.Generator().manual_seed(1000 + i)
        x = torch.randn(*shape, generator=gen)
        k = _kernel(spec, gen)
        y_ref = ref_upfirdn2d(x, k, up=up, down=down, pad=pad)
        y_or = so.upfirdn2d_ref(x, k, up=up, down=down, pad=pad)
        _close(y_or, y_ref, 1e-6, "upfirdn2d/" + name)
        out[name + ".x"] = x
        out[name + ".k"] = k
        out[name + ".cfg"] = np.array([up, down, pad[0], pad[1]])
        out[name + ".y"] = y_ref
    _save("upfirdn2d", **out)


def gen_fused_act(ref_models):
    from models.stylegan2.op.fused_act import fused_leaky_relu as ref_flr
    out = {}
    shapes = [(3, 5, 4, 4), (4, 7), (2, 6, 3, 5), (1, 4, 33, 31), (2, 8, 16, 16)]
    for i, shape in enumerate(shapes):
        gen = torch.Generator().manual_seed(2000 + i)
        x = torch.randn(*shape, generator=gen, requires_grad=True)
        b = torch.randn(shape[1], generator=gen, requires_grad=True)
        y_ref = ref_flr(x, b)  # CPU branch: leaky_relu(x + b, 0.2) * sqrt(2)
        g = torch.randn(*shape, generator=gen)
        gx_ref, gb_ref = torch.autograd.grad(y_ref, [x, b], g)
        y_or = so.fused_leaky_relu_ref(x.detach(), b.detach())
        gx_or, gb_or = so.fused_leaky_relu_backward_ref(g, y_or)
        _close(y_or, y_ref.detach(), 1e-6, "fused_leaky_relu fwd %s" % (shape,))
        _close(gx_or, gx_ref, 1e-6, "fused_leaky_relu gx %s" % (shape,))
        _close(gb_or, gb_ref, 1e-5, "fused_leaky_relu gb %s" % (shape,))
        tag = "case%d" % i
        out[tag + ".x"], out[tag + ".b"], out[tag + ".g"] = x.detach(), b.detach(), g
        out[tag + ".y"], out[tag + ".gx"], out[tag + ".gb"] = y_ref.detach(), gx_ref, gb_ref
    _save("fused_act", **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_models = refimport.import_reference()
    gen_upfirdn2d(ref_models)
    gen_fused_act(ref_models)
    for extra in EXTRA_GENERATORS:
        extra(ref_models)


def _smooth_grid(gen, n, res, theta, jitter):
    import torch.nn.functional as F
    base = F.affine_grid(theta, (n, 3, res, res), align_corners=False)
    coarse = torch.randn(n, 2, 5, 5, generator=gen)
    bump = F.interpolate(coarse, size=(res, res), mode="bicubic", align_corners=False).permute(0, 2, 3, 1)
    return (base + jitter * bump).contiguous()


WARP_CASES = [
    # name, source size, output res, affine rows per sample, jitter, padding mode
    ("minify_border", 32, 16, [[2.1, 0.4, 0.1, -0.4, 2.1, -0.2], [1.3, 0.0, 0.3, 0.0, 1.3, 0.0]], 0.00, "border"),
    ("minify_reflect", 32, 16, [[2.1, 0.4, 0.1, -0.4, 2.1, -0.2], [3.3, 0.0, 0.3, 0.0, 3.3, 0.5]], 0.00, "reflection"),
    ("minify_zeros", 32, 16, [[2.1, 0.4, 0.1, -0.4, 2.1, -0.2], [1.3, 0.0, 0.3, 0.0, 1.3, 0.0]], 0.00, "zeros"),
    ("flow_border", 64, 64, [[1.0, 0.0, 0.0, 0.0, 1.0, 0.0], [1.6, 0.2, 0.0, -0.2, 1.6, 0.1]], 0.15, "border"),
    ("flow_reflect", 64, 32, [[1.2, 0.0, 0.4, 0.0, 1.2, -0.3], [2.6, 0.2, 0.0, -0.2, 2.6, 0.1]], 0.20, "reflection"),
    ("nonpow2_src", 20, 12, [[1.9, 0.1, 0.0, -0.1, 1.9, 0.0], [0.7, 0.0, 0.0, 0.0, 0.7, 0.1]], 0.05, "border"),
    ("nonpow2_52", 52, 24, [[1.0, 0.0, 0.0, 0.0, 1.0, 0.0], [1.5, 0.5, 0.0, -0.5, 1.5, 0.0]], 0.02, "border"),
    ("identity", 16, 16, [[1.0, 0.0, 0.0, 0.0, 1.0, 0.0], [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]], 0.00, "border"),
]


def gen_mipmap_warp(ref_models):
    import torch.nn.functional as F
    from models.spatial_transformers.antialiased_sampling import MipmapWarp, Warp, BilinearDownsample
    from oracle import sampling as S
    out = {}
    for i, (name, size, res, rows, jitter, mode) in enumerate(WARP_CASES):
        gen = torch.Generator().manual_seed(3000 + i)
        x = torch.randn(2, 3, size, size, generator=gen, requires_grad=True)
        theta = torch.tensor(rows).reshape(2, 2, 3)
        grid = _smooth_grid(gen, 2, res, theta, jitter).requires_grad_(True)
        go = torch.randn(2, 3, res, res, generator=gen)
        mw = MipmapWarp(3.5)
        y_ref = mw(x, grid, padding_mode=mode)
        gx_ref, gg_ref = torch.autograd.grad(y_ref, [x, grid], go)
        levels_ref = mw.levels_map * 2.5
        # oracle restatement (explicit index arithmetic) against the reference, forward and autograd
        xo, go_ = x.detach().clone().requires_grad_(True), grid.detach().clone().requires_grad_(True)
        y_or, aux = S.mipmap_warp_ref(xo, go_, 3.5, 0.0, mode, return_aux=True)
        gx_or, gg_or = torch.autograd.grad(y_or, [xo, go_], go)
        _close(y_or, y_ref.detach(), 5e-6, "mipmap_warp/" + name)
        _close(gx_or, gx_ref, 2e-5, "mipmap_warp gx/" + name)
        _close(gg_or, gg_ref, 2e-4, "mipmap_warp ggrid/" + name)
        assert torch.equal(aux["levels"], levels_ref.detach()) or (aux["levels"] - levels_ref).abs().max() < 1e-6
        out[name + ".x"], out[name + ".grid"], out[name + ".go"] = x.detach(), grid.detach(), go
        out[name + ".mode"] = np.array(S.PAD_MODES.index(mode))
        out[name + ".y"], out[name + ".gx"], out[name + ".ggrid"] = y_ref.detach(), gx_ref, gg_ref
        out[name + ".levels"] = aux["levels"]
        # plain Warp on the same inputs
        yw = Warp()(x, grid, padding_mode=mode)
        gxw, ggw = torch.autograd.grad(yw, [x, grid], go)
        _close(S.warp_ref(x.detach(), grid.detach(), mode), yw.detach(), 5e-6, "warp/" + name)
        out[name + ".warp_y"], out[name + ".warp_gx"], out[name + ".warp_ggrid"] = yw.detach(), gxw, ggw
    _save("mipmap_warp", **out)
    # BilinearDownsample (resize_fake2stn, train.py:62)
    gen = torch.Generator().manual_seed(3100)
    x = torch.randn(2, 3, 32, 32, generator=gen)
    bd = {}
    for stride in (2, 4):
        y = BilinearDownsample(stride, 3)(x)
        _close(S.bilinear_downsample_ref(x, stride), y, 1e-6, "bilinear_downsample")
        bd["s%d.y" % stride] = y
    bd["x"] = x
    _save("bilinear_downsample", **bd)


def gen_flow(ref_models):
    from models.spatial_transformers import warping_heads as wh
    from oracle import flow as FL
    out = {}

    class _Head:  # FlowHead.__init__ calls .cuda() (warping_heads.py:158); its methods only need this attribute
        flow_downsample = 8

    for i, (n, h, w, s) in enumerate([(2, 4, 4, 8), (3, 6, 5, 4), (1, 8, 8, 8)]):
        gen = torch.Generator().manual_seed(4000 + i)
        _Head.flow_downsample = s
        low = (0.05 * torch.randn(n, h, w, 2, generator=gen)).requires_grad_(True)
        mask = torch.randn(n, 9 * s * s, h, w, generator=gen, requires_grad=True)
        base = (torch.eye(2, 3)[None] + 0.2 * torch.randn(n, 2, 3, generator=gen)).requires_grad_(True)
        ident = FL.identity_flow_ref(s * h, s * w)
        delta_ref = wh.FlowHead.upsample_flow(_Head, low, mask)
        flow_ref = wh.apply_affine(base, ident + delta_ref)
        gd = torch.randn(delta_ref.shape, generator=gen)
        gf = torch.randn(flow_ref.shape, generator=gen)
        grads_ref = torch.autograd.grad((delta_ref * gd).sum() + (flow_ref * gf).sum(), [low, mask, base])
        lo, mo, bo = [t.detach().clone().requires_grad_(True) for t in (low, mask, base)]
        delta_or, flow_or = FL.flow_compose_ref(lo, mo, ident, bo, None, s)
        grads_or = torch.autograd.grad((delta_or * gd).sum() + (flow_or * gf).sum(), [lo, mo, bo])
        _close(delta_or, delta_ref.detach(), 1e-6, "upsample_flow %d" % i)
        _close(flow_or, flow_ref.detach(), 1e-6, "apply_affine %d" % i)
        for a, b, nm in zip(grads_or, grads_ref, ("low", "mask", "base")):
            _close(a, b, 1e-5, "flow grads %s %d" % (nm, i))
        tag = "case%d" % i
        out[tag + ".low"], out[tag + ".mask"], out[tag + ".base"] = low.detach(), mask.detach(), base.detach()
        out[tag + ".s"] = np.array(s)
        out[tag + ".gd"], out[tag + ".gf"] = gd, gf
        out[tag + ".delta"], out[tag + ".flow"] = delta_ref.detach(), flow_ref.detach()
        out[tag + ".g_low"], out[tag + ".g_mask"], out[tag + ".g_base"] = grads_ref
    # similarity matrices (SimilarityHead.make_affine_matrix / make_3x3)
    gen = torch.Generator().manual_seed(4100)
    params = torch.randn(5, 8, generator=gen)  # K = 2
    m_ref = wh.SimilarityHead.make_affine_matrix(*torch.split(params, 2, dim=1))
    _close(FL.similarity_matrix_ref(params), m_ref, 1e-6, "similarity matrix")
    base = torch.randn(5, 2, 2, 3, generator=gen)
    one_hot = torch.tensor([0, 0, 1], dtype=torch.float).view(1, 1, 1, 3).expand(5, 2, 1, 3)
    comp_ref = base @ torch.cat([m_ref, one_hot], 2)
    _close(FL.compose_similarity_ref(base, m_ref), comp_ref, 1e-6, "similarity compose")
    out["sim.params"], out["sim.matrix"], out["sim.base"], out["sim.composed"] = params, m_ref, base, comp_ref
    _save("flow_compose", **out)


def gen_networks(ref_models):
    """End-to-end fixtures: reference Generator / STN forward on CPU with construction-order-independent seeded
    weights (oracle.opset.fill_parameters), explicit noise and fixed inputs."""
    from oracle import opset
    torch.Tensor.cuda = lambda self, *a, **k: self  # reference FlowHead.__init__ calls .cuda() (warping_heads.py:158)
    from models.stylegan2.networks import Generator
    from models.spatial_transformers.spatial_transformer import get_stn
    out = {}
    gen = torch.Generator().manual_seed(5000)
    g = opset.fill_parameters(Generator(32, 32, 2, channel_multiplier=2).eval(), 1)
    z = torch.randn(3, 32, generator=gen)
    noise = [torch.randn(3, 1, n.shape[2], n.shape[3], generator=gen) for n in g.make_noise(1)]
    with torch.no_grad():
        img, lat = g([z], noise=noise, return_latents=True)
    out["gen.z"], out["gen.image"], out["gen.latent"] = z, img, lat
    for i, n in enumerate(noise):
        out["gen.noise%d" % i] = n
    for transforms in (["similarity"], ["similarity", "flow"]):
        stn = get_stn(list(transforms), flow_size=64, supersize=128, channel_multiplier=0.5, num_heads=1).eval()
        opset.fill_parameters(stn, 3, gain=0.3)
        x = torch.randn(2, 3, 128, 128, generator=gen)
        with torch.no_grad():
            o, grid, fm = stn(x, return_warp=True, return_flow=True, padding_mode="reflection")
        tag = "stn_" + "_".join(transforms)
        out[tag + ".x"], out[tag + ".out"], out[tag + ".grid"], out[tag + ".fm"] = x, o, grid, fm
    # BASELINE config 1: similarity-only STN @64 on CPU, non-identity head bias, 3 padding modes
    stn = get_stn(["similarity"], flow_size=64, supersize=64, channel_multiplier=0.5, num_heads=1).eval()
    opset.fill_parameters(stn, 5, gain=0.3)
    x = torch.randn(4, 3, 64, 64, generator=gen)
    with torch.no_grad():
        stn.warp_head.linear.bias.copy_(torch.tensor([0.3, 0.2, 0.1, -0.1]))
        for mode in ("border", "reflection", "zeros"):
            o, grid, m = stn(x, return_warp=True, return_flow=True, padding_mode=mode)
            out["cfg1.out." + mode] = o
        out["cfg1.x"], out["cfg1.grid"], out["cfg1.M"] = x, grid, m
    _save("networks", **out)


def _mse(a, b):
    return (a - b).pow(2).mean(dim=(1, 2, 3))


def gen_losses(ref_models):
    """End-to-end pin of the CALLERS of the hot path: the reference's gangealing_loss (config 2 shape, shrunk) and
    gangealing_cluster_loss (config 5: K heads, flips, sample_from_full_res) on CPU with seeded weights; the RNG is the
    global CPU generator, consumed in the same order by this repo's mirror."""
    from oracle import opset
    torch.Tensor.cuda = lambda self, *a, **k: self
    from models.stylegan2.networks import Generator
    from models.spatial_transformers.spatial_transformer import get_stn
    from models.spatial_transformers.antialiased_sampling import BilinearDownsample
    from models.latent_learner import DirectionInterpolator
    from models.losses.loss import gangealing_loss, gangealing_cluster_loss, total_variation_loss
    out = {}
    for tag, heads, flips, full_res in (("uni", 1, False, False), ("cluster", 2, True, True)):
        gen_size = 128 if full_res else 64
        # DirectionInterpolator's random buffers are hard-coded 512-d (latent_learner.py:39-40): 512-d latent
        g = opset.fill_parameters(Generator(gen_size, 512, 2, channel_multiplier=1).eval(), 11)
        for prm in g.parameters():
            prm.requires_grad = False
        stn = get_stn(["similarity", "flow"], flow_size=64, supersize=gen_size, channel_multiplier=0.25, num_heads=heads)
        opset.fill_parameters(stn, 12, gain=0.2)
        ll = DirectionInterpolator(None, 2, 3, g.n_latent, num_heads=heads)
        opset.fill_parameters(ll, 13, gain=0.5)
        resize = BilinearDownsample(2, 3) if full_res else torch.nn.Sequential()
        torch.manual_seed(1234)
        if heads == 1:
            loss, delta = gangealing_loss(g, stn, ll, _mse, resize, 0.6, 2, 512, False, "cpu", sample_from_full_res=full_res,
                                          padding_mode="reflection")
        else:
            loss, delta = gangealing_cluster_loss(g, stn, ll, _mse, resize, 0.6, 2, 512, False, heads, flips, "cpu",
                                                  sample_from_full_res=full_res, padding_mode="reflection")
        tv = total_variation_loss(delta)
        names = [n for n, _ in stn.named_parameters()]
        grads = torch.autograd.grad(loss + 10.0 * tv, list(stn.parameters()) + [ll.coefficients], allow_unused=True)
        out[tag + ".loss"], out[tag + ".tv"], out[tag + ".delta"] = loss.detach(), tv.detach(), delta.detach()
        picked = 0
        for n, gr in zip(names + ["ll.coefficients"], grads):
            if gr is not None and gr.abs().max() > 0 and picked < 6 and (gr.numel() < 5000 or n == "ll.coefficients"):
                out[tag + ".grad." + n] = gr
                picked += 1
    _save("losses", **out)


def gen_perceptual(ref_models):
    """Front end of the perceptual loss: the reference's normalize_tensor / NetLinLayer /
    spatial_average (models/losses/lpips.py:26-28, :197-205, :226, :238-247) on seeded relu-like feature maps."""
    import models.losses.lpips as L
    g = torch.Generator().manual_seed(77)
    out = {}
    for name, (n, c, h, w), weighted in (("c64", (2, 64, 8, 8), False), ("c128_w", (2, 128, 5, 7), True),
                                         ("c512", (1, 512, 3, 4), False), ("c256_w", (3, 256, 4, 4), True),
                                         ("c8", (2, 8, 6, 6), False)):
        f0 = torch.relu(torch.randn(n, c, h, w, generator=g)).requires_grad_(True)
        f1 = torch.relu(torch.randn(n, c, h, w, generator=g) + 0.3).requires_grad_(True)
        with torch.no_grad():
            f0[0, :, 0, 0] = 0.0            # an all-zero pixel in ONE map: |f| = 0 -> f/(0+eps) = 0 in the forward
        diff = (L.normalize_tensor(f0) - L.normalize_tensor(f1)) ** 2
        if weighted:
            lin = L.NetLinLayer(c, use_dropout=False)
            wv = torch.rand(c, generator=g)
            with torch.no_grad():
                lin.model[-1].weight.copy_(wv.reshape(1, c, 1, 1))
            res = L.spatial_average(lin(diff), keepdim=True)
            out[name + ".weight"] = wv
        else:
            res = L.spatial_average(diff.sum(dim=1, keepdim=True), keepdim=True)
        go = torch.randn(res.shape, generator=g)
        g0, g1 = torch.autograd.grad(res, [f0, f1], go)
        # the reference's sqrt backward is nan at an all-zero pixel (inf * 0); the fixture stores the finite remainder
        out[name + ".f0"], out[name + ".f1"], out[name + ".out"], out[name + ".gout"] = f0.detach(), f1.detach(), res.detach(), go
        out[name + ".g0"], out[name + ".g1"] = torch.nan_to_num(g0, nan=0.0), g1
    _save("perceptual", **out)


def gen_points(ref_models):
    """Point transfer: the reference's congeal_points / uncongeal_points /
    transfer_points (spatial_transformer.py:631-726) for a similarity-only and a composed similarity+flow STN on CPU.
    congeal_points on a flow STN is integer work (argmin + unravel_index): stored for exact comparison."""
    from oracle import opset
    torch.Tensor.cuda = lambda self, *a, **k: self
    from models.spatial_transformers.spatial_transformer import get_stn
    gen = torch.Generator().manual_seed(6100)
    out = {}
    for transforms in (["similarity"], ["similarity", "flow"]):
        tag = "pts_" + "_".join(transforms)
        stn = get_stn(list(transforms), flow_size=64, supersize=64, channel_multiplier=0.25, num_heads=1).eval()
        opset.fill_parameters(stn, 21, gain=0.3)
        img_a = torch.randn(2, 3, 64, 64, generator=gen)
        img_b = torch.randn(2, 3, 64, 64, generator=gen)
        pts = torch.rand(2, 9, 2, generator=gen) * 40.0 + 12.0          # pixel coordinates well inside the image
        with torch.no_grad():
            congealed = stn.congeal_points(img_a, pts)
            back = stn.uncongeal_points(img_b, congealed.float() if congealed.dtype != torch.float32 else congealed,
                                        normalize_input_points=congealed.dtype != torch.float32)
            moved = stn.transfer_points(img_a, img_b, pts)
        out[tag + ".img_a"], out[tag + ".img_b"], out[tag + ".points"] = img_a, img_b, pts
        out[tag + ".congealed"], out[tag + ".uncongealed"], out[tag + ".transferred"] = congealed, back, moved
    _save("points", **out)


STN_OPTION_CASES = [
    # name, transforms, heads, call kwargs (tensors are created by the generator / the test from the same seeds)
    ("iters3", ["similarity"], 1, dict(iters=3, return_warp=True, return_flow=True, padding_mode="border")),
    ("composed_iters2_alpha", ["similarity", "flow"], 1, dict(iters=2, alpha=[0.35, 0.8], return_warp=True, return_flow=True,
                                                              return_sim=True, padding_mode="reflection")),
    ("composed_outres", ["similarity", "flow"], 1, dict(output_resolution=96, return_warp=True, return_flow=True,
                                                         padding_mode="zeros")),
    ("heads2_cartesian", ["similarity", "flow"], 2, dict(return_warp=True, return_flow=True, padding_mode="border")),
    ("heads2_unfold", ["similarity", "flow"], 2, dict(unfold=True, return_warp=True, return_flow=True, padding_mode="border")),
    ("intermediates", ["similarity", "flow"], 1, dict(return_intermediates=True, padding_mode="border")),
]


def stn_option_kwargs(kw):
    """Case-table kwargs -> call kwargs (per-sample alpha is a tensor, warping_heads.py:244)."""
    kw = dict(kw)
    if isinstance(kw.get("alpha"), list):
        kw["alpha"] = torch.tensor(kw["alpha"])
    return kw


def _flatten_outputs(res):
    """STN return values (tensor, list/tuple of tensors, nested) -> flat list of tensors in traversal order."""
    if torch.is_tensor(res):
        return [res]
    out = []
    for r in res:
        out += _flatten_outputs(r)
    return out


def gen_stn_options(ref_models):
    """Orchestration options of the STN API: iterated similarity, composed similarity+flow with
    alpha / output_resolution / return_sim / return_intermediates, multi-head cartesian policy and unfold --
    reference spatial_transformer.py:78-139, :523-615 on CPU with seeded weights."""
    from oracle import opset
    torch.Tensor.cuda = lambda self, *a, **k: self
    from models.spatial_transformers.spatial_transformer import get_stn
    out = {}
    for i, (name, transforms, heads, kw) in enumerate(STN_OPTION_CASES):
        gen = torch.Generator().manual_seed(7000 + i)
        stn = get_stn(list(transforms), flow_size=64, supersize=64, channel_multiplier=0.25, num_heads=heads).eval()
        opset.fill_parameters(stn, 31 + i, gain=0.3)
        x = torch.randn(2 if "alpha" in kw else 1, 3, 64, 64, generator=gen)
        with torch.no_grad():
            res = stn(x, **stn_option_kwargs(kw))
        out["opt_" + name + ".x"] = x
        for j, t in enumerate(_flatten_outputs(res)):
            out["opt_%s.out%d" % (name, j)] = t
    _save("stn_options", **out)


def gen_perceptual_loss(ref_models):
    """The whole perceptual loss as the training script builds it (lpips.py:13-17: LPIPS(net='vgg', lpips=False,
    pnet_rand=True) / 18) with seeded VGG16 weights: value and input gradients on small images.  The weights are
    regenerated from the seed by the test (oracle.opset.fill_convs_in_order), only inputs/outputs are stored."""
    import models.losses.lpips as L
    from oracle import opset
    net = L.LPIPS(net="vgg", lpips=False, pnet_rand=True, verbose=False)
    opset.fill_convs_in_order(net, 4242)
    g = torch.Generator().manual_seed(4243)
    in0 = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).requires_grad_(True)
    in1 = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).requires_grad_(True)
    val = net(in0, in1) / 18.0
    g0, g1 = torch.autograd.grad(val.sum(), [in0, in1])
    _save("perceptual_loss", in0=in0.detach(), in1=in1.detach(), val=val.detach(), g0=g0, g1=g1)


def classifier_setup(mods, heads=2, flips=True):
    """The models of one cluster-classifier iteration (shrunk config 5), built from the module classes in `mods` (the
    reference's here, this repo's in tests/test_classifier_cpu.py) with the same seeded weights."""
    from oracle import opset
    gen_size, clusters = 128, heads * (1 + int(flips))
    g = opset.fill_parameters(mods["Generator"](gen_size, 512, 2, channel_multiplier=1).eval(), 41)
    stn = mods["get_stn"](["similarity", "flow"], flow_size=64, supersize=gen_size, channel_multiplier=0.25, num_heads=heads)
    opset.fill_parameters(stn, 42, gain=0.2)
    ll = mods["DirectionInterpolator"](None, 2, 3, g.n_latent, num_heads=heads)
    opset.fill_parameters(ll, 43, gain=0.5)
    cls = mods["ResnetClassifier"](64, channel_multiplier=0.25, num_heads=clusters, supersize=gen_size)
    opset.fill_parameters(cls, 44, gain=1.0)     # unit gain: logits of order 1, assignments differ between samples
    resize = mods["BilinearDownsample"](2, 3)
    for m in (g, stn, ll):
        for prm in m.parameters():
            prm.requires_grad = False
    return g, stn, ll, cls, resize, clusters


def classifier_decimate(images):
    """Image outputs of the run_* helpers are stored as an asymmetric 16x16 sub-grid (a mirrored image lands on other pixels)."""
    return images[..., ::8, 3::8]


def gen_classifier(ref_models):
    """BASELINE config 5, second half: the reference's ResnetClassifier (models/cluster_classifier.py) -- logits, the
    run_* inference helpers (exact index outputs), `accuracy` -- and one iteration of train_cluster_classifier.py:84-105
    (assignments by the frozen clustering STN, cross-entropy, gradients, one Adam step) on CPU with seeded weights."""
    from torch import nn, optim
    torch.Tensor.cuda = lambda self, *a, **k: self
    from models import ResnetClassifier, accuracy
    from models.stylegan2.networks import Generator
    from models.spatial_transformers.spatial_transformer import get_stn
    from models.spatial_transformers.antialiased_sampling import BilinearDownsample
    from models.latent_learner import DirectionInterpolator
    from models.losses.loss import assign_fake_images_to_clusters
    mods = dict(Generator=Generator, get_stn=get_stn, DirectionInterpolator=DirectionInterpolator,
                ResnetClassifier=ResnetClassifier, BilinearDownsample=BilinearDownsample)
    g, stn, ll, cls, resize, clusters = classifier_setup(mods)
    out = {}
    gen = torch.Generator().manual_seed(4400)
    x = torch.randn(6, 3, 128, 128, generator=gen)          # larger than stn_in_size: goes through input_downsample
    x = x * torch.linspace(0.3, 2.0, 6).reshape(6, 1, 1, 1) + torch.linspace(-1, 1, 6).reshape(6, 1, 1, 1)
    dec = classifier_decimate
    with torch.no_grad():
        out["cls.x"], out["cls.logits"] = x, cls(x)
        out["cls.assign"], out["cls.assign_noflip"] = cls.assign(x), cls.assign(x, ignore_flips=True)
        for c in range(clusters // 2):
            kept, preds, flip, keep = cls.run(x, c, return_flip_indices=True)
            out["cls.run%d.kept" % c], out["cls.run%d.preds" % c] = dec(kept), preds
            out["cls.run%d.flip" % c], out["cls.run%d.keep" % c] = flip, keep
            flipped, flip_t = cls.run_flip_target(x, c)
            out["cls.run_flip_target%d.out" % c], out["cls.run_flip_target%d.flip" % c] = dec(flipped), flip_t
        flipped, preds, classes, flip = cls.run_flip(x)
        out["cls.run_flip.out"], out["cls.run_flip.classes"], out["cls.run_flip.flip"] = dec(flipped), classes, flip
        tiled, policy = cls.run_flip_cartesian(x)
        out["cls.cartesian.out"], out["cls.cartesian.policy"] = dec(tiled), policy
    pr, gt = torch.randn(16, clusters, generator=gen), torch.randn(16, clusters, generator=gen)
    out["acc.pred"], out["acc.gt"] = pr, gt
    out["acc.k1"], out["acc.k2"], out["acc.k3"] = accuracy(pr, gt), accuracy(pr, gt, k=2), accuracy(pr, gt, k=3)
    # one training iteration, statements of train_cluster_classifier.py:84-105
    batch, psi = 3, 0.0
    xent = nn.CrossEntropyLoss()
    cls_optim = optim.Adam(cls.parameters(), lr=0.001)
    torch.manual_seed(4321)
    with torch.no_grad():
        assigned, _, _, _, resized, distance = assign_fake_images_to_clusters(
            g, stn, ll, _mse, resize, psi, batch, 512, True, 2, True, "cpu", sample_from_full_res=True, z=None,
            padding_mode="reflection")
    predicted = cls(resized[:batch])
    loss = xent(predicted, assigned.indices)
    out["step.xent"], out["step.acc1"], out["step.acc2"] = loss.detach(), accuracy(predicted, -distance), accuracy(predicted, -distance, k=2)
    out["step.assignments"], out["step.distance"], out["step.logits"] = assigned.indices, distance, predicted.detach()
    out["step.hist_gt"] = torch.bincount(assigned.indices, minlength=clusters).div(float(batch))
    out["step.hist_pred"] = torch.bincount(predicted.argmax(dim=1), minlength=clusters).div(float(batch))
    cls.zero_grad()
    loss.backward()
    picked = 0
    for n, prm in cls.named_parameters():
        if prm.grad is not None and prm.grad.abs().max() > 0 and prm.numel() < 5000 and picked < 6:
            out["step.grad." + n] = prm.grad.clone()
            picked += 1
    cls_optim.step()
    out["step.after.to_logits.bias"] = cls.to_logits.bias.detach().clone()
    out["step.after.final_conv.1.bias"] = cls.final_conv[1].bias.detach().clone()
    _save("classifier", **out)


EXTRA_GENERATORS = [gen_mipmap_warp, gen_flow, gen_networks, gen_losses, gen_perceptual, gen_points, gen_stn_options,
                    gen_perceptual_loss, gen_classifier]

if __name__ == "__main__":
    main()
