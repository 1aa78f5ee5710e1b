"""GPU: the host-side networks on the sm_100a op set against the reference-generated fixtures and against the same
networks run on the CPU oracle (forward AND gradients, identical weights / latents / noise)."""
import zlib

import pytest
import torch

from conftest import assert_close, load_golden
from oracle import opset

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_generator_on_gpu_matches_reference_fixture():
    from gangealing_b200.stylegan2 import Generator
    blob = load_golden("networks")
    g = opset.fill_parameters(Generator(32, 32, 2, channel_multiplier=2).eval(), 1).to(DEV)
    noise = [blob["gen.noise%d" % i].to(DEV) for i in range(g.num_layers)]
    with torch.no_grad():
        img, lat = g([blob["gen.z"].to(DEV)], noise=noise, return_latents=True)
    assert_close(lat, blob["gen.latent"], rtol=1e-4, what="latent")
    assert_close(img, blob["gen.image"], rtol=1e-3, what="image")   # north-star tolerance, cuDNN TF32 convs included


@pytest.mark.parametrize("transforms", [("similarity",), ("similarity", "flow")])
def test_stn_on_gpu_matches_reference_fixture(transforms):
    from gangealing_b200.stn import get_stn
    blob = load_golden("networks")
    tag = "stn_" + "_".join(transforms)
    stn = get_stn(list(transforms), flow_size=64, supersize=128, channel_multiplier=0.5, num_heads=1).eval()
    opset.fill_parameters(stn, 3, gain=0.3).to(DEV)
    with torch.no_grad():
        out, grid, fm = stn(blob[tag + ".x"].to(DEV), return_warp=True, return_flow=True, padding_mode="reflection")
    assert_close(grid, blob[tag + ".grid"], rtol=1e-3, what="grid")
    assert_close(fm, blob[tag + ".fm"], rtol=1e-3, what="flow/matrix")
    assert_close(out, blob[tag + ".out"], rtol=2e-3, what="warped image")


def test_full_size_generator_fp32_exact_convs_vs_cpu_oracle():
    """256^2 generator, batch 1, TF32 off: every hand-written op on the path vs the CPU restatement end to end."""
    from gangealing_b200.stylegan2 import Generator
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        g_cpu = Generator(256, 64, 2, channel_multiplier=2, ops=opset.cpu_ops()).eval()
        g_gpu = Generator(256, 64, 2, channel_multiplier=2).eval()
        g_gpu.load_state_dict(g_cpu.state_dict())
        g_gpu.to(DEV)
        z = torch.randn(1, 64)
        noise = g_cpu.make_noise(1)
        with torch.no_grad():
            a, _ = g_cpu([z], noise=noise)
            b, _ = g_gpu([z.to(DEV)], noise=[n.to(DEV) for n in noise])
        assert_close(b, a, rtol=1e-3, what="G(256) image")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_train_step_gradients_match_cpu_oracle():
    """Same weights, latents and noise: loss and STN / latent-learner gradients, GPU op set vs CPU oracle."""
    from gangealing_b200.training import TrainConfig, Trainer
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = TrainConfig(gen_size=64, flow_size=64, dim_latent=32, n_mlp=2, batch=2, inject=3, tv_weight=10.0)
        t_cpu = Trainer(cfg, "cpu", ops=opset.cpu_ops())
        t_gpu = Trainer(cfg, DEV)
        for a, b in ((t_cpu.generator, t_gpu.generator), (t_cpu.t_module, t_gpu.t_module), (t_cpu.ll_module, t_gpu.ll_module),
                     (t_cpu.loss_fn, t_gpu.loss_fn)):
            b.load_state_dict(a.state_dict())
        # make the heads non-trivial (they are zero-initialised) and freeze the noise
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():
            for tr in (t_cpu, t_gpu):
                for name, prm in tr.t_module.named_parameters():
                    if "warp_head" in name:
                        g.manual_seed(zlib.crc32(name.encode()) % 1000)   # str hashes are salted per process
                        prm.copy_((0.05 * torch.randn(prm.shape, generator=g)).to(prm.device))
        noise = t_cpu.generator.make_noise(cfg.batch)
        z = torch.randn(cfg.batch, cfg.dim_latent, generator=g)

        def run(tr, dev):
            import gangealing_b200.stylegan2.networks as nets
            it = {"i": 0}
            fixed = [n.to(dev) for n in noise] * 2
            orig = nets.NoiseInjection.sample

            def sample(batch, h, w, like):   # deterministic noise: one tensor per StyledConv call, in call order
                cands = [n for n in fixed if n.shape[2] == h and n.shape[3] == w]
                it["i"] += 1
                return cands[it["i"] % len(cands)]
            nets.NoiseInjection.sample = staticmethod(sample)
            try:
                ld = tr.losses(z.to(dev))
                full = ld["p"] + cfg.tv_weight * ld["tv"]
                grads = torch.autograd.grad(full, list(tr.t_module.parameters()) + [tr.ll_module.coefficients], allow_unused=True)
            finally:
                nets.NoiseInjection.sample = orig
            return ld, grads

        ld_c, g_c = run(t_cpu, "cpu")
        ld_g, g_g = run(t_gpu, DEV)
        assert_close(ld_g["p"], ld_c["p"], rtol=2e-3, what="perceptual loss")
        assert_close(ld_g["tv"], ld_c["tv"], rtol=2e-3, what="tv loss")
        names = [n for n, _ in t_cpu.t_module.named_parameters()] + ["ll.coefficients"]
        checked = 0
        for n, a, b in zip(names, g_c, g_g):
            if a is None or b is None:
                assert a is None and b is None, n
                continue
            if a.abs().max() < 1e-7:
                continue
            # long fp32 chains through two networks (and discontinuous LOD selection in the sampler): tiny-magnitude
            # tensors are judged loosely on their own scale, the gradient as a whole tightly on the global scale
            assert_close(b, a, rtol=6e-2, what="grad " + n)
            checked += 1
        assert checked > 20
        pairs = [(a, b) for a, b in zip(g_c, g_g) if a is not None and b is not None]
        assert_close(torch.cat([b.flatten().cpu() for _, b in pairs]), torch.cat([a.flatten() for a, _ in pairs]),
                     rtol=2e-3, what="whole gradient")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_config4_point_transfer_and_splat_vs_cpu_oracle():
    """BASELINE config 4, shrunk: flow STN at a higher output resolution -> uncongeal_points -> splat_points
    (applications/propagate_to_images.py:44-78), GPU op set vs the same host code on the CPU oracle."""
    from gangealing_b200.splat2d import splat2d
    from gangealing_b200.stn import get_stn
    from oracle import splat as SP
    kw = dict(flow_size=64, supersize=128, channel_multiplier=0.25, num_heads=1)
    s_cpu = get_stn(["similarity", "flow"], ops=opset.cpu_ops(), **kw).eval()
    opset.fill_parameters(s_cpu, 21, gain=0.2)
    s_gpu = get_stn(["similarity", "flow"], **kw).eval()
    s_gpu.load_state_dict(s_cpu.state_dict())
    s_gpu.to(DEV)
    g = torch.Generator().manual_seed(3)
    imgs = (torch.rand(2, 3, 128, 128, generator=g) * 2 - 1)
    ys, xs = torch.meshgrid(torch.arange(64.), torch.arange(64.), indexing="ij")
    disc = ((ys - 32) ** 2 + (xs - 32) ** 2) < (0.35 * 64) ** 2
    pts = torch.stack([xs[disc], ys[disc]], dim=1)[None].repeat(2, 1, 1)          # congealed-frame pixel coordinates
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            pc = s_cpu.uncongeal_points(imgs, pts, normalize_input_points=True, output_resolution=128, padding_mode="border")
            pg = s_gpu.uncongeal_points(imgs.to(DEV), pts.to(DEV), normalize_input_points=True, output_resolution=128,
                                        padding_mode="border")
        assert_close(pg, pc, atol=2e-2, what="transferred points (pixels)")
        colors = torch.randn(2, pts.shape[1], 3, generator=g)
        expect = SP.splat_points_ref(imgs, pc, 1.3, 0.75, colors)
        got = SP.splat_points_ref(imgs.to(DEV), pg, 1.3, 0.75, colors.to(DEV), splat_fn=splat2d)
        assert_close(got, expect, rtol=5e-3, what="propagated image")
    finally:
        torch.backends.cudnn.allow_tf32 = old


def _patched_noise(noise_by_size):
    """Context: NoiseInjection.sample / the fused path's sampler return fixed tensors (per size, in call order)."""
    import contextlib
    import gangealing_b200.stylegan2.networks as nets

    @contextlib.contextmanager
    def ctx(dev):
        it = {"i": 0}
        orig = nets.NoiseInjection.sample

        def sample(batch, h, w, like):     # both the layer-by-layer and the fused synthesis draw through this hook
            cands = noise_by_size[(h, w)]
            it["i"] += 1
            return cands[it["i"] % len(cands)].to(dev)[:batch]
        nets.NoiseInjection.sample = staticmethod(sample)
        try:
            yield
        finally:
            nets.NoiseInjection.sample = orig
    return ctx


def test_cluster_config5_step_matches_cpu_oracle():
    """BASELINE config 5 (K heads, flips, sample_from_full_res, reflection padding; reference loss.py:78-92), shrunk: the
    clustering loss on the GPU op set vs the same host code on the CPU oracle -- loss, EXACT cluster assignments, the
    assigned heads' residual flow and the gradients."""
    from gangealing_b200.training import TrainConfig, Trainer
    from gangealing_b200.training.losses import assign_fake_images_to_clusters
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = TrainConfig(gen_size=128, flow_size=64, dim_latent=64, n_mlp=2, batch=3, inject=3, num_heads=4, flips=True,
                          ndirs=2, sample_from_full_res=True, padding_mode="reflection", gen_channel_multiplier=1,
                          stn_channel_multiplier=0.25, tv_weight=10.0)
        t_cpu = Trainer(cfg, "cpu", ops=opset.cpu_ops())
        t_gpu = Trainer(cfg, DEV)
        for a, b in ((t_cpu.generator, t_gpu.generator), (t_cpu.t_module, t_gpu.t_module), (t_cpu.ll_module, t_gpu.ll_module),
                     (t_cpu.loss_fn, t_gpu.loss_fn)):
            b.load_state_dict(a.state_dict())
        g = torch.Generator().manual_seed(9)
        with torch.no_grad():
            for tr in (t_cpu, t_gpu):
                for name, prm in list(tr.t_module.named_parameters()) + list(tr.ll_module.named_parameters()):
                    if "warp_head" in name or "coefficients" in name:
                        g.manual_seed(zlib.crc32(name.encode()) % 1000)
                        prm.copy_((0.08 * torch.randn(prm.shape, generator=g)).to(prm.device))
        sizes = sorted({(n.shape[2], n.shape[3]) for n in t_cpu.generator.make_noise(1)})
        noise = {hw: [torch.randn(4 * cfg.batch, 1, hw[0], hw[1], generator=g) for _ in range(3)] for hw in sizes}
        z = torch.randn(cfg.batch, cfg.dim_latent, generator=g)
        res = []
        for tr, dev in ((t_cpu, "cpu"), (t_gpu, DEV)):
            with _patched_noise(noise)(dev):
                assign, _, delta, _, _, collapsed = assign_fake_images_to_clusters(
                    tr.generator, tr.stn, tr.ll, tr.loss_fn, tr.resize_fake2stn, tr.psi_t, cfg.batch, cfg.dim_latent,
                    cfg.freeze_ll, cfg.num_heads, cfg.flips, dev, sample_from_full_res=True, z=z.to(dev), padding_mode="reflection")
            with _patched_noise(noise)(dev):
                ld = tr.losses(z.to(dev))
                full = ld["p"] + cfg.tv_weight * ld["tv"]
                grads = torch.autograd.grad(full, list(tr.t_module.parameters()) + [tr.ll_module.coefficients], allow_unused=True)
            res.append((assign, collapsed, ld, grads))
        (a_c, col_c, ld_c, g_c), (a_g, col_g, ld_g, g_g) = res
        assert_close(col_g, col_c, rtol=2e-3, what="per-(image, head, flip) perceptual scores")
        # assignments are integer work: exact wherever the two best scores are separated by more than the float tolerance
        top2 = col_c.topk(2, dim=1, largest=False).values
        decided = (top2[:, 1] - top2[:, 0]) > 5e-3 * top2[:, 1]
        assert decided.any()
        assert torch.equal(a_g.indices.cpu()[decided], a_c.indices[decided])
        assert_close(ld_g["p"], ld_c["p"], rtol=2e-3, what="cluster perceptual loss")
        assert_close(ld_g["tv"], ld_c["tv"], rtol=3e-3, what="tv of the assigned heads' flows")
        pairs = [(a, b) for a, b in zip(g_c, g_g) if a is not None and b is not None]
        assert len(pairs) > 20
        assert_close(torch.cat([b.flatten().cpu() for _, b in pairs]), torch.cat([a.flatten() for a, _ in pairs]),
                     rtol=8e-3, what="whole gradient")   # min over 8 (head, flip) scores of two chained networks: fp32 chains
        out = t_gpu.step()
        assert all(torch.isfinite(v) for v in out.values())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_config4_full_size_point_transfer_and_splat_vs_cpu_oracle():
    """BASELINE config 4 at its real size: flow STN with supersize 512 and output_resolution 512, P = 25 233 points (disc of
    radius 0.35*R rendered at R = 256, applications/propagate_to_images.py:44-78) -> uncongeal_points -> splat at sigma 0.3
    and 1.3, one frame, GPU op set vs the CPU oracle."""
    from gangealing_b200.splat2d import splat2d
    from gangealing_b200.stn import get_stn
    from oracle import splat as SP
    kw = dict(flow_size=128, supersize=512, channel_multiplier=0.125, num_heads=1)
    s_cpu = get_stn(["similarity", "flow"], ops=opset.cpu_ops(), **kw).eval()
    opset.fill_parameters(s_cpu, 31, gain=0.2)
    s_gpu = get_stn(["similarity", "flow"], **kw).eval()
    s_gpu.load_state_dict(s_cpu.state_dict())
    s_gpu.to(DEV)
    g = torch.Generator().manual_seed(4)
    img = (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1)
    R = 256
    ys, xs = torch.meshgrid(torch.arange(float(R)), torch.arange(float(R)), indexing="ij")
    disc = ((ys - R / 2) ** 2 + (xs - R / 2) ** 2) < (0.35 * R) ** 2
    pts = torch.stack([xs[disc], ys[disc]], dim=1)[None] * (127.0 / (R - 1))     # congealed-frame (128 px) coordinates
    assert pts.shape[1] == 25233
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            pc = s_cpu.uncongeal_points(img, pts, normalize_input_points=True, output_resolution=512, padding_mode="border")
            pg = s_gpu.uncongeal_points(img.to(DEV), pts.to(DEV), normalize_input_points=True, output_resolution=512,
                                        padding_mode="border")
        assert_close(pg, pc, atol=5e-2, what="transferred points (pixels, 512^2 frame)")
        colors = torch.randn(1, pts.shape[1], 3, generator=g)
        for sigma in (0.3, 1.3):
            expect = SP.splat_points_ref(img, pc, sigma, 0.75, colors)
            got = SP.splat_points_ref(img.to(DEV), pc.to(DEV), sigma, 0.75, colors.to(DEV), splat_fn=splat2d)
            assert_close(got, expect, rtol=1e-3, what="propagated image, sigma %.1f" % sigma)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_bf16_training_step_tracks_the_fp32_step():
    """BASELINE config 3 (bf16 activations, fp32 master weights / accumulation) vs config 2 (fp32) on identical weights,
    latents and noise: the loss within bf16 rounding noise, the gradient direction preserved."""
    from gangealing_b200.training import TrainConfig, Trainer
    kw = dict(gen_size=128, flow_size=64, dim_latent=64, n_mlp=2, batch=4, inject=3, tv_weight=100.0, gen_channel_multiplier=1,
              stn_channel_multiplier=0.5)
    t32 = Trainer(TrainConfig(dtype="f32", **kw), DEV)
    t16 = Trainer(TrainConfig(dtype="bf16", **kw), DEV)
    for a, b in ((t32.t_module, t16.t_module), (t32.ll_module, t16.ll_module)):
        b.load_state_dict(a.state_dict())
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for tr in (t32, t16):
            for name, prm in tr.t_module.named_parameters():
                if "warp_head" in name:
                    g.manual_seed(zlib.crc32(name.encode()) % 1000)
                    prm.copy_((0.05 * torch.randn(prm.shape, generator=g)).to(prm.device))
    z = torch.randn(4, 64, generator=g).to(DEV)
    res = []
    for tr in (t32, t16):
        assert tr.generator.act_dtype == (torch.float32 if tr is t32 else torch.bfloat16)
        torch.manual_seed(123)                        # identical device-side noise draws (fp32 noise in both modes)
        ld = tr.losses(z)
        full = ld["p"] + tr.cfg.tv_weight * ld["tv"]
        grads = torch.autograd.grad(full, list(tr.t_module.parameters()) + [tr.ll_module.coefficients], allow_unused=True)
        # keep plain numbers only: a live autograd graph would pin its AccumulateGrad nodes to the (legacy) stream of this
        # eager call and break the graph capture below
        res.append(({k: float(v) for k, v in ld.items()}, torch.cat([x.flatten().float() for x in grads if x is not None]).clone()))
        del ld, full, grads
    (l32, g32), (l16, g16) = res
    assert abs(float(l16["p"]) - float(l32["p"])) <= 3e-2 * abs(float(l32["p"])), (float(l16["p"]), float(l32["p"]))
    assert abs(float(l16["tv"]) - float(l32["tv"])) <= 5e-2 * abs(float(l32["tv"])) + 1e-8
    cos = torch.nn.functional.cosine_similarity(g16, g32, dim=0).item()
    assert cos > 0.98, cos
    assert all(p.dtype == torch.float32 for p in t16.t_module.parameters())   # fp32 master weights
    out = t16.step()
    assert all(torch.isfinite(v) for v in out.values())
    t16.capture(warmup=2)                             # the bf16 step is graph-capturable like the fp32 one
    out = t16.step()
    torch.cuda.synchronize()
    assert all(torch.isfinite(v) for v in out.values())


@pytest.mark.parametrize("transforms", [("similarity",), ("similarity", "flow")])
def test_point_transfer_on_gpu_matches_reference_fixture(transforms):
    """congeal_points / uncongeal_points / transfer_points on the sm_100a op set (the flow STN's nearest-neighbour search
    runs in the tiled argmin kernel, csrc/points.cu) against the reference-generated fixture.  The nearest-neighbour indices
    are integer work: equal to the reference's up to the GPU convolutions' effect on the grid itself (an index may move by
    one cell where two cells are equidistant to within the float noise of the network)."""
    from gangealing_b200.stn import get_stn
    blob = load_golden("points")
    tag = "pts_" + "_".join(transforms)
    stn = get_stn(list(transforms), flow_size=64, supersize=64, channel_multiplier=0.25, num_heads=1).eval()
    opset.fill_parameters(stn, 21, gain=0.3).to(DEV)
    img_a, img_b, pts = blob[tag + ".img_a"].to(DEV), blob[tag + ".img_b"].to(DEV), blob[tag + ".points"].to(DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            congealed = stn.congeal_points(img_a, pts)
            is_index = congealed.dtype != torch.float32
            ref = blob[tag + ".congealed"]
            back = stn.uncongeal_points(img_b, ref.to(DEV).float() if is_index else congealed, normalize_input_points=is_index)
            moved = stn.transfer_points(img_a, img_b, pts)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert congealed.dtype == ref.dtype and congealed.shape == ref.shape
    if is_index:
        diff = (congealed.cpu() - ref).abs()
        assert diff.max() <= 1 and (diff == 0).float().mean() >= 0.8, diff
    else:
        assert_close(congealed, ref, rtol=1e-4, what="congealed points")
    assert_close(back, blob[tag + ".uncongealed"], rtol=2e-4, what="uncongealed points")
    assert_close(moved, blob[tag + ".transferred"], atol=1.0 if is_index else 1e-2, what="transferred points (pixels)")


def test_uncongeal_and_splat_equals_the_two_step_path():
    """propagate_to_images' per-frame work (uncongeal_points -> splat_points) with the lookup fused into the splat vs the
    reference's two steps through the same STN."""
    from gangealing_b200.stn import get_stn
    from gangealing_b200.splat2d import splat2d
    from oracle import splat as SP
    stn = get_stn(["similarity", "flow"], flow_size=64, supersize=128, channel_multiplier=0.25, num_heads=1).eval()
    opset.fill_parameters(stn, 21, gain=0.2).to(DEV)
    g = torch.Generator().manual_seed(3)
    imgs = (torch.rand(2, 3, 128, 128, generator=g) * 2 - 1).to(DEV)
    ys, xs = torch.meshgrid(torch.arange(64.), torch.arange(64.), indexing="ij")
    disc = ((ys - 32) ** 2 + (xs - 32) ** 2) < (0.35 * 64) ** 2
    pts = torch.stack([xs[disc], ys[disc]], dim=1)[None].repeat(2, 1, 1).to(DEV)
    colors = torch.randn(2, pts.shape[1], 3, generator=g).to(DEV)
    with torch.no_grad():
        fused_img, fused_pts = stn.uncongeal_and_splat(imgs, pts, colors, 1.3, 0.75, output_resolution=128,
                                                       normalize_input_points=True, padding_mode="border")
        two_pts = stn.uncongeal_points(imgs, pts, normalize_input_points=True, output_resolution=128, padding_mode="border")
        two_img = SP.splat_points_ref(imgs, two_pts, 1.3, 0.75, colors, splat_fn=splat2d)
    assert_close(fused_pts, two_pts, atol=2e-3, what="points")
    assert_close(fused_img, two_img, rtol=2e-3, what="propagated image")
