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


def test_cluster_config5_step_runs_on_gpu():
    """BASELINE config 5 shape (K heads, flips, sample_from_full_res, reflection padding), shrunk: one optimisation step."""
    from gangealing_b200.training import TrainConfig, Trainer
    cfg = TrainConfig(gen_size=128, flow_size=64, dim_latent=64, n_mlp=2, batch=2, inject=3, num_heads=2, flips=True,
                      ndirs=2, sample_from_full_res=True, padding_mode="reflection", gen_channel_multiplier=1,
                      stn_channel_multiplier=0.25)
    tr = Trainer(cfg, DEV)
    out1 = tr.step()
    out2 = tr.step()
    assert all(torch.isfinite(v) for v in out2.values()) and float(out2["p"]) > 0
