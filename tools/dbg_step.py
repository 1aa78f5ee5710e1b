"""Per-parameter gradient error of one train step, GPU op set vs CPU oracle, channels_last off/on (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import opset
from gangealing_b200.training import TrainConfig, Trainer
import gangealing_b200.stylegan2.networks as nets
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
def build(cl, dev, ops=None):
    cfg = TrainConfig(gen_size=64, flow_size=64, dim_latent=32, n_mlp=2, batch=2, inject=3, tv_weight=10.0, channels_last=cl)
    return cfg, Trainer(cfg, dev, ops=ops)
cfg, t_cpu = build(False, "cpu", opset.cpu_ops())
g = torch.Generator().manual_seed(5)
noise = t_cpu.generator.make_noise(cfg.batch)
z = torch.randn(cfg.batch, cfg.dim_latent, generator=g)
def prep(tr):
    with torch.no_grad():
        for name, prm in tr.t_module.named_parameters():
            if "warp_head" in name:
                g.manual_seed(__import__('zlib').crc32(name.encode()) % 1000)
                prm.copy_((0.05 * torch.randn(prm.shape, generator=g)).to(prm.device))
def run(tr, dev):
    it = {"i": 0}
    fixed = [n.to(dev) for n in noise] * 2
    orig = nets.NoiseInjection.sample
    def sample(batch, h, w, like):
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
prep(t_cpu)
ld_c, g_c = run(t_cpu, "cpu")
names = [n for n, _ in t_cpu.t_module.named_parameters()] + ["ll.coefficients"]
res = {}
for cl in (False, True):
    _, t_gpu = build(cl, DEV)
    for a, b in ((t_cpu.generator, t_gpu.generator), (t_cpu.t_module, t_gpu.t_module), (t_cpu.ll_module, t_gpu.ll_module),
                 (t_cpu.loss_fn, t_gpu.loss_fn)):
        b.load_state_dict(a.state_dict())
    ld_g, g_g = run(t_gpu, DEV)
    print("channels_last", cl, "loss p", ld_g["p"].item(), ld_c["p"].item(), "tv", ld_g["tv"].item(), ld_c["tv"].item())
    res[cl] = g_g
    rows = []
    for n, a, b in zip(names, g_c, g_g):
        if a is None or b is None: continue
        sc = a.abs().max().item()
        if sc < 1e-7: continue
        rows.append(((b.cpu() - a).abs().max().item() / sc, sc, n))
    rows.sort(reverse=True)
    for r in rows[:8]:
        print("   rel %.3e  mag %.3e  %s" % r)
rows = []
for n, a, b in zip(names, res[False], res[True]):
    if a is None or b is None: continue
    sc = a.abs().max().item()
    if sc < 1e-7: continue
    rows.append(((b - a).abs().max().item() / sc, sc, n))
rows.sort(reverse=True)
print("NHWC vs NCHW on GPU:")
for r in rows[:8]:
    print("   rel %.3e  mag %.3e  %s" % r)
