"""CPU: checkpoints in the reference's layout (train.py:20-27 / :214-224) -- the optimiser-dict split / merge helpers, the
host logic of FusedAdamEMA.load_state_dict (its kernel is not launched here: the constructor's CUDA requirements are
patched out) and Trainer.checkpoint() / load_checkpoint() on the oracle op set."""
import copy

import pytest
import torch

from oracle import opset


def _params(gen, shapes, channels_last=False):
    out = []
    for s in shapes:
        t = torch.randn(*s, generator=gen)
        if channels_last and t.dim() == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        out.append(t.requires_grad_(True))
    return out


def _step(optimisers, params, gen, n=1):
    for _ in range(n):
        for p in params:
            p.grad = torch.randn(p.shape, generator=gen)
        for o in optimisers:
            o.step()


def test_split_and_merge_follow_the_reference_checkpoint_layout():
    from gangealing_b200.training.fused_optim import merge_adam_state_dicts, split_adam_state_dict
    gen = torch.Generator().manual_seed(0)
    a, b = _params(gen, [(4, 3, 3, 3), (7,), (5, 2)]), _params(gen, [(6,), (2, 9)])
    opt_a, opt_b = torch.optim.Adam(a, lr=1e-3), torch.optim.Adam(b, lr=1e-2)     # the reference's t_optim / ll_optim
    _step([opt_a, opt_b], a + b, gen, 3)
    merged = merge_adam_state_dicts([opt_a.state_dict(), opt_b.state_dict()])
    assert [g["params"] for g in merged["param_groups"]] == [[0, 1, 2], [3, 4]]
    assert [g["lr"] for g in merged["param_groups"]] == [1e-3, 1e-2]
    # one optimiser with two groups (the fused step's layout) continues exactly where the two left off
    a2, b2 = [p.detach().clone().requires_grad_(True) for p in a], [p.detach().clone().requires_grad_(True) for p in b]
    both = torch.optim.Adam([{"params": a2, "lr": 1.0}, {"params": b2, "lr": 1.0}])
    both.load_state_dict(copy.deepcopy(merged))    # (Optimizer.load_state_dict aliases same-dtype tensors of an in-memory dict)
    g1, g2 = torch.Generator().manual_seed(9), torch.Generator().manual_seed(9)
    _step([opt_a, opt_b], a + b, g1, 2)
    _step([both], a2 + b2, g2, 2)
    for p, q in zip(a + b, a2 + b2):
        assert torch.equal(p, q)
    parts = split_adam_state_dict(both.state_dict())
    assert len(parts) == 2 and parts[1]["param_groups"][0]["params"] == [0, 1]
    for part, ref in zip(parts, (opt_a, opt_b)):
        want = ref.state_dict()
        assert part["param_groups"][0]["lr"] == want["param_groups"][0]["lr"]
        assert sorted(part["state"]) == sorted(want["state"])
        for i in want["state"]:
            assert torch.equal(part["state"][i]["exp_avg"], want["state"][i]["exp_avg"])
            assert float(part["state"][i]["step"]) == float(want["state"][i]["step"]) == 5.0


@pytest.fixture
def fused_on_cpu(monkeypatch):
    """FusedAdamEMA's bookkeeping without a GPU: the CUDA-tensor check and the pinned staging buffer are patched out; `step()`
    (the kernel launch) is never called."""
    from gangealing_b200 import _lib
    monkeypatch.setattr(_lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    from gangealing_b200.training.fused_optim import FusedAdamEMA
    return FusedAdamEMA


def test_fused_adam_load_state_dict_keeps_addresses_and_the_shared_step(fused_on_cpu):
    from gangealing_b200.training.fused_optim import merge_adam_state_dicts, split_adam_state_dict
    gen = torch.Generator().manual_seed(1)
    shapes_a, shapes_b = [(8, 4, 3, 3), (8,), (8, 4, 1, 1)], [(5,), (3, 6)]
    ref_a, ref_b = _params(gen, shapes_a), _params(gen, shapes_b)                 # reference side: NCHW-strided everything
    opt_a, opt_b = torch.optim.Adam(ref_a, lr=2e-3), torch.optim.Adam(ref_b, lr=3e-2)
    _step([opt_a, opt_b], ref_a + ref_b, gen, 4)
    ours_a = [p.detach().clone(memory_format=torch.channels_last if p.dim() == 4 else torch.contiguous_format).requires_grad_(True)
              for p in ref_a]                                                     # the Trainer stores 4-D weights channels-last
    ours_b = [p.detach().clone().requires_grad_(True) for p in ref_b]
    opt = fused_on_cpu([{"params": ours_a, "lr": 1e-3}, {"params": ours_b, "lr": 1e-2}])
    before = {p: (opt.state[p]["exp_avg"].data_ptr(), opt.state[p]["exp_avg_sq"].data_ptr()) for p in ours_a + ours_b}
    opt.load_state_dict(copy.deepcopy(merge_adam_state_dicts([opt_a.state_dict(), opt_b.state_dict()])))
    assert float(opt._state3[0]) == 4.0
    assert [float(t) for t in opt._lr] == pytest.approx([2e-3, 3e-2])
    for p, r, ro in zip(ours_a + ours_b, ref_a + ref_b, [opt_a] * 3 + [opt_b] * 2):
        st = opt.state[p]
        assert (st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()) == before[p], "a moment tensor moved"
        assert st["exp_avg"].stride() == p.stride()
        assert torch.equal(st["exp_avg"], ro.state[r]["exp_avg"]) and torch.equal(st["exp_avg_sq"], ro.state[r]["exp_avg_sq"])
        assert st["step"].data_ptr() == opt._state3.data_ptr() and float(st["step"]) == 4.0
    # and back out: the reference's optimisers accept what this one saves
    t_sd, ll_sd = split_adam_state_dict(opt.state_dict())
    fresh_a, fresh_b = torch.optim.Adam(_params(gen, shapes_a), lr=1.0), torch.optim.Adam(_params(gen, shapes_b), lr=1.0)
    fresh_a.load_state_dict(copy.deepcopy(t_sd))
    fresh_b.load_state_dict(copy.deepcopy(ll_sd))
    assert fresh_a.param_groups[0]["lr"] == pytest.approx(2e-3) and fresh_b.param_groups[0]["lr"] == pytest.approx(3e-2)
    for fp, r in zip(fresh_a.param_groups[0]["params"], ref_a):
        assert torch.equal(fresh_a.state[fp]["exp_avg"], opt_a.state[r]["exp_avg"])
        assert float(fresh_a.state[fp]["step"]) == 4.0
    # a checkpoint whose two optimisers disagree on the step count cannot feed ONE fused counter
    _step([opt_b], ref_b, gen, 1)
    with pytest.raises(RuntimeError, match="step count"):
        opt.load_state_dict(merge_adam_state_dicts([opt_a.state_dict(), opt_b.state_dict()]))
    # a checkpoint taken before the first step: moments zeroed, counter reset
    opt.load_state_dict(merge_adam_state_dicts([torch.optim.Adam(ref_a, lr=5e-4).state_dict(), torch.optim.Adam(ref_b, lr=5e-3).state_dict()]))
    assert float(opt._state3[0]) == 0.0 and all(float(opt.state[p]["exp_avg"].abs().max()) == 0.0 for p in ours_a + ours_b)


def test_trainer_checkpoint_round_trip_resumes_identically():
    from gangealing_b200.training import TrainConfig, Trainer
    cpu = opset.cpu_ops()
    cfg = TrainConfig(gen_size=64, flow_size=64, dim_latent=16, n_mlp=1, batch=1, inject=3, stn_channel_multiplier=0.25,
                      gen_channel_multiplier=1)
    a = Trainer(cfg, "cpu", ops=cpu)
    for _ in range(2):
        a.step()
    ckpt = copy.deepcopy(a.checkpoint(iteration=2))
    assert {"g_ema", "t", "t_ema", "t_optim", "ll", "ll_optim"} <= set(ckpt)
    assert float(ckpt["t_optim"]["state"][0]["step"]) == 2.0
    ckpt.update(t_sched={"last_epoch": 2}, ll_sched={"last_epoch": 2}, args={"note": "keys of the reference's file that have no use here"})
    b = Trainer(cfg, "cpu", ops=cpu)      # same seed = same frozen perceptual network (not part of a checkpoint) ...
    b.step()                              # ... but its own training state until the checkpoint is loaded
    with torch.no_grad():
        for prm in list(b.t_module.parameters()) + list(b.t_ema.parameters()) + list(b.generator.parameters()):
            prm.add_(0.01)
    assert b.load_checkpoint(ckpt) is True                       # also moves b to iteration 2 of the schedule (psi, learning rates)
    a.set_iteration(2)
    assert float(a.psi_t) == float(b.psi_t)
    z = torch.randn(cfg.batch, cfg.dim_latent)
    for tr in (a, b):
        torch.manual_seed(77)                                                # same noise draws in both
        tr.step(z)
    for (n, p), (_, q) in zip(a.t_module.named_parameters(), b.t_module.named_parameters()):
        assert torch.equal(p, q), n
    for p, q in zip(a.t_ema.parameters(), b.t_ema.parameters()):
        assert torch.equal(p, q)
    assert torch.equal(a.ll_module.coefficients, b.ll_module.coefficients)
    # generator-only restore (train.py --load_G_only / a StyleGAN2 checkpoint without GANgealing state)
    import dataclasses
    c = Trainer(dataclasses.replace(cfg, seed=12), "cpu", ops=cpu)
    assert c.load_checkpoint({"g_ema": ckpt["g_ema"]}) is False
    assert all(torch.equal(p, q) for p, q in zip(a.generator.parameters(), c.generator.parameters()))
    assert not all(torch.equal(p, q) for p, q in zip(a.t_module.parameters(), c.t_module.parameters()))


def test_classifier_trainer_checkpoint_round_trip():
    from gangealing_b200.training import ClassifierTrainer, TrainConfig, Trainer
    cpu = opset.cpu_ops()
    cfg = TrainConfig(gen_size=64, flow_size=64, dim_latent=16, n_mlp=1, batch=2, inject=3, num_heads=2, flips=True, ndirs=2,
                      stn_channel_multiplier=0.25, gen_channel_multiplier=1, padding_mode="reflection")
    a = ClassifierTrainer(Trainer(cfg, "cpu", ops=cpu), ops=cpu)
    for _ in range(2):
        a.step()
    ckpt = copy.deepcopy(a.checkpoint())
    assert {"classifier", "g_ema", "t_ema", "ll", "cls_optim"} <= set(ckpt)      # train_cluster_classifier.py:25-29
    b = ClassifierTrainer(Trainer(cfg, "cpu", ops=cpu), ops=cpu)
    b.step()
    assert b.load_checkpoint(ckpt) is True
    z = torch.randn(cfg.batch, cfg.dim_latent)
    for tr in (a, b):
        torch.manual_seed(5)
        out = tr.step(z)
    assert all(torch.equal(p, q) for p, q in zip(a.module.parameters(), b.module.parameters()))
    assert torch.isfinite(out["cross_entropy"])
    # a GANgealing checkpoint without a classifier (the first launch of train_cluster_classifier.py): trunk from the STN
    c = ClassifierTrainer(Trainer(cfg, "cpu", ops=cpu), ops=cpu, init_from_stn=False)
    assert c.load_checkpoint({k: ckpt[k] for k in ("g_ema", "t_ema", "ll")}) is False
    trunk = dict(c.trainer.t_ema.stns[0].named_parameters())
    assert all(torch.equal(p, trunk[n]) for n, p in c.module.named_parameters() if n.startswith(("convs.", "final_conv.")))


def test_eager_step_is_refused_while_a_captured_graph_is_held():
    """The captured graph re-reads the multi-tensor kernels' pinned pointer tables; an eager step would overwrite them."""
    from gangealing_b200.training import TrainConfig, Trainer
    cfg = TrainConfig(gen_size=64, flow_size=64, dim_latent=16, n_mlp=1, batch=1, inject=3, stn_channel_multiplier=0.25,
                      gen_channel_multiplier=1)
    tr = Trainer(cfg, "cpu", ops=opset.cpu_ops())
    tr._graph = object()                      # stands in for a torch.cuda.CUDAGraph (none can exist on CPU)
    with pytest.raises(RuntimeError, match="release_graph"):
        tr._eager_step()
    tr.release_graph()
    assert tr._graph is None and torch.isfinite(tr.step()["p"])
