"""GPU: the fused training-loop bookkeeping (csrc/optim.cu) against its references --
FusedAdamEMA vs torch.optim.Adam (the reference's optimiser, train.py:204-205) + the reference's `accumulate`
(models/__init__.py:19-24) evaluated on the CPU, and the fused total-variation loss vs models/losses/loss.py:4-12."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _reference_tv(delta_flow):     # models/losses/loss.py:4-12, reduce_batch=True
    def dist(a):
        return torch.where(a <= 1.0, 0.5 * a.pow(2), a - 0.5).mean()
    dy = dist((delta_flow[:, :-1] - delta_flow[:, 1:]).abs())
    dx = dist((delta_flow[:, :, :-1] - delta_flow[:, :, 1:]).abs())
    return dx + dy


@pytest.mark.parametrize("shape,scale", [((3, 16, 16, 2), 0.3), ((2, 128, 128, 2), 1.5), ((1, 7, 33, 2), 4.0), ((2, 2, 2, 2), 1.0)])
def test_total_variation_loss_kernel(shape, scale):
    from gangealing_b200.stn.transformer import total_variation_loss
    g = torch.Generator().manual_seed(shape[1])
    f = torch.randn(*shape, generator=g) * scale       # differences on both sides of the Huber knee
    fo = f.clone().requires_grad_(True)
    lo = _reference_tv(fo)
    (go,) = torch.autograd.grad(lo, fo, torch.tensor(2.5))
    fg = f.to(DEV).requires_grad_(True)
    lg = total_variation_loss(fg)
    assert lg.shape == ()
    assert_close(lg, lo, rtol=1e-5, what="tv loss")
    (gg,) = torch.autograd.grad(lg, fg, torch.tensor(2.5, device=DEV))
    assert_close(gg, go, rtol=1e-5, what="tv gradient")
    # per-sample form (forward_with_flip's tie-break) keeps the tensor formulation and agrees with it
    per = total_variation_loss(f.to(DEV), reduce_batch=False)
    assert per.shape == (shape[0],)


def test_fused_adam_ema_matches_torch_adam_and_the_reference_ema():
    from gangealing_b200.training.fused_optim import FusedAdamEMA
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 32, 3, 3), (130,), (7, 5), (256, 64, 1, 1), (3, 70001), (1,)]
    ref_a = [torch.randn(*s, generator=g).requires_grad_(True) for s in shapes[:4]]
    ref_b = [torch.randn(*s, generator=g).requires_grad_(True) for s in shapes[4:]]
    ref_ema = [p.detach().clone() for p in ref_a]
    opt_a = torch.optim.Adam(ref_a, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    opt_b = torch.optim.Adam(ref_b, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    decay = 0.5 ** (32 / 10000)

    def dev(p, cl):
        t = p.detach().to(DEV)
        if cl and t.dim() == 4:
            t = t.contiguous(memory_format=torch.channels_last)     # the Trainer stores 4-D STN weights channels-last
        return t.requires_grad_(True)
    our_a = [dev(p, True) for p in ref_a]
    our_b = [dev(p, False) for p in ref_b]
    our_ema = [p.detach().clone(memory_format=torch.preserve_format) for p in our_a]
    opt = FusedAdamEMA([{"params": our_a, "lr": 1e-3}, {"params": our_b, "lr": 1e-2}], ema_pairs=dict(zip(our_a, our_ema)),
                       ema_decay=decay)
    for it in range(6):
        if it == 3:          # learning-rate schedule: device scalars on our side
            for grp in opt_a.param_groups:
                grp["lr"] = 4e-4
            opt.set_lr(0, 4e-4)
            opt.set_lr(1, torch.tensor(2e-3, device=DEV))
            for grp in opt_b.param_groups:
                grp["lr"] = 2e-3
        for p, q in zip(ref_a + ref_b, our_a + our_b):
            gr = torch.randn(p.shape, generator=g) * (10.0 ** (it - 3))
            p.grad = gr.clone()
            q.grad = gr.to(DEV).contiguous(memory_format=torch.channels_last) if (q.dim() == 4 and q.is_contiguous(memory_format=torch.channels_last)) else gr.to(DEV)
        opt_a.step(); opt_b.step()
        with torch.no_grad():
            for e, p in zip(ref_ema, ref_a):
                e.mul_(decay).add_(p.data, alpha=1 - decay)          # models/__init__.py:23-24
        opt.step()
    for p, q in zip(ref_a + ref_b, our_a + our_b):
        assert_close(q, p, rtol=2e-6, what="parameter %s" % (tuple(p.shape),))
    for e, q in zip(ref_ema, our_ema):
        assert_close(q, e, rtol=2e-6, what="ema")
    # optimiser state carries torch.optim.Adam's keys and values
    st_ref, st_our = opt_a.state[ref_a[0]], opt.state[our_a[0]]
    assert set(st_our.keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(st_our["step"]) == float(st_ref["step"]) == 6.0
    assert_close(st_our["exp_avg"], st_ref["exp_avg"], rtol=2e-6, what="exp_avg")
    assert_close(st_our["exp_avg_sq"], st_ref["exp_avg_sq"], rtol=2e-6, what="exp_avg_sq")
    sd = opt.state_dict()
    assert len(sd["param_groups"]) == 2 and len(sd["state"]) == len(shapes)


def test_fused_adam_ema_is_graph_capturable():
    from gangealing_b200.training.fused_optim import FusedAdamEMA
    p = torch.randn(1000, device=DEV).requires_grad_(True)
    e = p.detach().clone()
    grad = torch.randn(1000, device=DEV)
    p.grad = grad
    opt = FusedAdamEMA([{"params": [p], "lr": 1e-3}], ema_pairs={p: e}, ema_decay=0.9)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt.step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    before = p.detach().clone()
    opt.set_lr(0, 0.0)              # a replay reads the device scalar: lr 0 leaves the parameter untouched
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(p.detach(), before)
    opt.set_lr(0, 1e-2)
    graph.replay()
    torch.cuda.synchronize()
    assert not torch.equal(p.detach(), before)
    assert float(opt.state[p]["step"]) == 3.0       # side-stream step + two replays (the capture itself executes nothing)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_multi_tensor_weight_scaling_equals_the_per_layer_products(dtype):
    """op/scaled_weights.WeightScaler (the STN's `weight * scale` products of reference networks.py:121-127,146-149 and their
    backward as a few multi-tensor launches) vs the per-layer ATen products: the same fp32 multiply and the same rounding, so
    after several optimiser steps from identical states the parameters must agree to cuDNN's run-to-run noise.  Also: the
    scaler really served the layers (it learns each layer's dtype on the first step and is active from the second)."""
    from gangealing_b200.training import TrainConfig, Trainer
    kw = dict(gen_size=64, flow_size=32, dim_latent=32, n_mlp=2, batch=2, inject=3, gen_channel_multiplier=1,
              stn_channel_multiplier=0.25, tv_weight=10.0, dtype=dtype)
    ta = Trainer(TrainConfig(fused_weight_scaling=True, **kw), DEV)
    tb = Trainer(TrainConfig(fused_weight_scaling=False, **kw), DEV)
    assert ta.weight_scaler is not None and tb.weight_scaler is None
    n_layers = len(ta.weight_scaler.by_module)
    assert n_layers >= 20 and len(ta.weight_scaler.groups) >= 1
    tb.t_module.load_state_dict(ta.t_module.state_dict())
    tb.ll_module.load_state_dict(ta.ll_module.state_dict())
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():    # leave the zero-initialised identity warp so that every gradient is non-trivial
        for tr in (ta, tb):
            g.manual_seed(5)
            for name, prm in tr.t_module.named_parameters():
                if "warp_head" in name:
                    prm.copy_((0.05 * torch.randn(prm.shape, generator=g)).to(DEV))
    zs = [torch.randn(2, 32, generator=g).to(DEV) for _ in range(3)]
    for tr in (ta, tb):
        for i, z in enumerate(zs):
            torch.manual_seed(100 + i)            # identical device-side noise draws
            tr.step(z)
    served = [e for e in ta.weight_scaler.by_module.values() if e.dtype is not None]
    assert len(served) == n_layers                                       # every layer was seen ...
    assert all(len(grp.tables) >= 2 for grp in ta.weight_scaler.groups)  # ... and every group launched forward AND backward
    pa = torch.cat([p.detach().flatten() for p in ta.t_module.parameters()])
    pb = torch.cat([p.detach().flatten() for p in tb.t_module.parameters()])
    assert torch.isfinite(pa).all()
    assert_close(pa, pb, rtol=2e-5 if dtype == "f32" else 2e-3, what="STN parameters after 3 steps")
    # the cache never outlives a step, and a captured graph replays the same launches
    assert all(grp.outputs is None for grp in ta.weight_scaler.groups) and not ta.weight_scaler.active
    ta.capture(warmup=2)
    out = ta.step(zs[0])
    assert all(torch.isfinite(v) for v in out.values())
