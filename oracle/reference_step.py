"""The UNMODIFIED reference's training iteration on the host CPU cores -- the `cpu_baseline` / `--impl reference` arm of
bench.py ("kind": "reference").  Test/bench infrastructure, never on the product path.

What runs is the reference's own code, byte-compiled from where it lies under /root/reference into oracle/_ref/refpy_cpu
by oracle/build_ref.py (the checkout itself does not exist on the GPU box): its Generator, get_stn, DirectionInterpolator,
BilinearDownsample, LPIPS(net='vgg', lpips=False, pnet_rand=True)/18, gangealing_loss, total_variation_loss and accumulate,
driven by the statements of train.py:106-136 (loss -> zero_grad -> backward -> Adam x2 -> EMA), with the optimisers of
train.py:204-205.  Two things are stubbed, as SURVEY.md 8(c) found necessary to execute it without a GPU:
  * torch.utils.cpp_extension.load (op/upfirdn2d.py:9-16, op/fused_act.py:10-17 JIT-build CUDA at import) -- the ops
    then take their native CPU branches (upfirdn2d.py:146-149, fused_act.py:87-94);
  * Tensor.cuda -> identity while the flow head is constructed (warping_heads.py:158 calls .cuda() in __init__).
Weights: seeded random initialisation (no checkpoints exist offline), the same recipe as BASELINE config 2.
"""
import contextlib
import os
import sys
import time

import torch

from . import build_ref


def available():
    return build_ref.refpy_cpu_available()


@contextlib.contextmanager
def _reference_modules():
    """Import context: the reference's `models` / `utils` packages from refpy_cpu, isolated from whatever the process
    has registered under those names (the compat shim of the GPU drop-in tests uses the same names)."""
    import torch.utils.cpp_extension as cpp_ext

    class _NoNative:
        def __getattr__(self, name):
            raise RuntimeError("reference native extension is stubbed out (CPU branches only)")

    def is_ref(k):
        return k in ("models", "utils") or k.startswith("models.") or k.startswith("utils.")

    saved = {k: v for k, v in sys.modules.items() if is_ref(k)}
    for k in saved:
        del sys.modules[k]
    old_load, old_cuda = cpp_ext.load, torch.Tensor.cuda
    cpp_ext.load = lambda *a, **k: _NoNative()
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, build_ref.REFPY_CPU)
    try:
        import models
        assert models.__file__.startswith(build_ref.REFPY_CPU), models.__file__
        yield models
    finally:
        sys.path.remove(build_ref.REFPY_CPU)
        cpp_ext.load, torch.Tensor.cuda = old_load, old_cuda
        for k in [k for k in sys.modules if is_ref(k)]:
            del sys.modules[k]
        sys.modules.update(saved)


class ReferenceStep:
    """BASELINE config 2 built from the reference's own constructors (train.py:197-205) on the CPU."""

    def __init__(self, batch, gen_size=256, flow_size=128, dim_latent=512, n_mlp=8, seed=0, threads=None):
        if not available():
            raise RuntimeError("oracle/_ref/refpy_cpu not built (python -m oracle.build_ref, needs /root/reference)")
        if threads:
            torch.set_num_threads(threads)
        self.batch, self.dim_latent = batch, dim_latent
        self._ctx = _reference_modules()
        m = self.m = self._ctx.__enter__()
        torch.manual_seed(seed)
        from torch import optim
        import models.losses.lpips as L
        self.generator = m.Generator(gen_size, dim_latent, n_mlp, channel_multiplier=2).eval()
        kw = dict(flow_size=flow_size, supersize=flow_size, channel_multiplier=0.5, num_heads=1)
        self.stn = m.get_stn(["similarity", "flow"], **kw)
        self.t_ema = m.get_stn(["similarity", "flow"], **kw)
        self.ll = m.DirectionInterpolator(pca_path=None, n_comps=1, inject_index=5, n_latent=self.generator.n_latent, num_heads=1)
        m.accumulate(self.t_ema, self.stn, 0)
        m.requires_grad(self.generator, False)
        net = L.LPIPS(net="vgg", lpips=False, pnet_rand=True, verbose=False)
        self.loss_fn = lambda x, y: net(x, y) / 18.0
        self.resize = m.BilinearDownsample(gen_size // flow_size, 3)
        self.t_optim = optim.Adam(self.stn.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
        self.ll_optim = optim.Adam(self.ll.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
        self.accum = 0.5 ** (32 / (10 * 1000))

    def close(self):
        if self._ctx is not None:
            self._ctx.__exit__(None, None, None)
            self._ctx = None

    def step(self, psi=0.5, tv_weight=2500.0):
        m = self.m
        p, delta = m.gangealing_loss(self.generator, self.stn, self.ll, self.loss_fn, self.resize, psi, self.batch,
                                     self.dim_latent, False, "cpu", sample_from_full_res=False, padding_mode="border")
        tv = m.total_variation_loss(delta)
        self.stn.zero_grad()
        self.ll.zero_grad()
        (p + tv_weight * tv).backward()
        self.t_optim.step()
        self.ll_optim.step()
        m.accumulate(self.t_ema, self.stn, self.accum)
        return {"p": p.detach(), "tv": tv.detach()}


def step_rate(batch, steps=1, warmup=0, threads=None):
    """-> (images/s, seconds per step, losses of the last step) of the reference iteration on the host cores."""
    ref = ReferenceStep(batch, threads=threads)
    try:
        for _ in range(warmup):
            ref.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = ref.step()
        dt = time.perf_counter() - t0
        return batch * steps / dt, dt / steps, {k: float(v) for k, v in out.items()}
    finally:
        ref.close()


if __name__ == "__main__":
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    print(step_rate(b, threads=os.cpu_count()))
