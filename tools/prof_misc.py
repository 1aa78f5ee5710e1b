"""Tiny driver for ncu: ONE launch (after one warm-up launch) of every hand-written kernel family that is NOT covered by
tools/prof_nhwc.py / prof_tail.py -- the STN's one-pass sampler (forward + backward), flow composition backward, the
perceptual front end (distance fwd/bwd, VGG slice boundary fwd/bwd), BilinearDownsample, the fused optimiser, the TV loss,
splat2d (scatter + normalise, with and without the fused lookup), the nearest-neighbour search, the x2 resamplers and the
tcgen05 demodulation GEMM -- at the shapes of the bench (per-GPU batch B, default 32) / BASELINE config 4.

  ncu --set full --clock-control none --import-source on -k regex:gg:: -o gpurun_out/misc python tools/prof_misc.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200 import op  # noqa: E402
from gangealing_b200.op import style_path  # noqa: E402
from gangealing_b200.op.feature_distance import feature_distance  # noqa: E402
from gangealing_b200.op.vgg_pool import bias_relu_pool  # noqa: E402
from gangealing_b200.splat2d import nn_argmin, splat2d, splat2d_lookup  # noqa: E402
from gangealing_b200.stn import sampling as S  # noqa: E402
from gangealing_b200.stn.transformer import total_variation_loss  # noqa: E402
from gangealing_b200.training.fused_optim import FusedAdamEMA  # noqa: E402

B = int(os.environ.get("B", "32"))
dev = "cuda"
CL = torch.channels_last
g = torch.Generator(device=dev).manual_seed(0)
k4 = torch.tensor([1., 3., 3., 1.]); k4 = k4[None] * k4[:, None]; k4 = (k4 / k4.sum() * 4).to(dev)


def rnd(*shape, scale=1.0):
    return torch.randn(*shape, device=dev, generator=g) * scale


# persistent inputs
img256 = rnd(B, 3, 256, 256).clamp(-1, 1)
img128 = rnd(B, 3, 128, 128).clamp(-1, 1)
theta = torch.eye(2, 3, device=dev)[None].repeat(B, 1, 1) + rnd(B, 2, 3, scale=0.05)
low, mask = rnd(B, 16, 16, 2, scale=0.02), rnd(B, 576, 16, 16)
ident = F.affine_grid(torch.eye(2, 3, device=dev)[None], (1, 1, 128, 128), align_corners=False)
f64 = [torch.relu(rnd(B, 64, 128, 128)).contiguous(memory_format=CL) for _ in range(2)]
f512 = [torch.relu(rnd(B, 512, 16, 16)).contiguous(memory_format=CL) for _ in range(2)]
raw64, b64 = rnd(B, 64, 128, 128).contiguous(memory_format=CL), rnd(64)
tent = torch.tensor([1., 3., 3., 1.], device=dev) / 8
kh, kv = tent[None, None, None, :].repeat(3, 1, 1, 1), tent[None, None, :, None].repeat(3, 1, 1, 1)
params = [rnd(43_000_000 // 8) for _ in range(8)]        # the STN's 43 M parameters as 8 tensors
for p in params:
    p.requires_grad_(True)
ema = {p: p.detach().clone() for p in params}
opt = FusedAdamEMA([{"params": params, "lr": 1e-3}], ema_pairs=ema, ema_decay=0.999)
flow_res = rnd(B, 128, 128, 2, scale=1.5)
R = 1024
ys, xs = torch.meshgrid(torch.arange(float(R)), torch.arange(float(R)), indexing="ij")
disc = ((ys - R / 2) ** 2 + (xs - R / 2) ** 2) < (0.35 * R) ** 2
pts = (torch.stack([xs[disc], ys[disc]], 1) * (511.0 / (R - 1)) + 0.25)[None].to(dev).contiguous()
vals = rnd(1, pts.shape[1], 3)
sig = torch.tensor([1.3], device=dev)
blank = torch.zeros(1, 3, 512, 512, device=dev)
grid512 = F.affine_grid(torch.eye(2, 3, device=dev)[None] * 0.9, (1, 1, 512, 512), align_corners=False)
query = (pts / 511.0) * 2 - 1
w_mod = [rnd(1, 512, 512, 3, 3), rnd(1, 256, 512, 3, 3)]
st_mod = [rnd(B, 512), rnd(B, 512)]

def sec_sampler():
    # STN one-pass sampler: similarity (256 -> 128, sample_from_full_res) and flow (128 -> 128), forward + backward
    src = img256.clone().requires_grad_(True)
    th = theta.clone().requires_grad_(True)
    out, grid, _ = S.stn_sample_affine(src, th, (128, 128), 3.5, 0.0, "border")
    out.square().mean().backward()
    src2, lo, mk = img128.clone().requires_grad_(True), low.clone().requires_grad_(True), mask.clone().requires_grad_(True)
    out, flow, delta, _ = S.stn_sample_flow(src2, lo, mk, ident, theta, None, 8, 3.5, 0.0, "border")
    (out.square().mean() + delta.square().mean()).backward()


def sec_perceptual():
    a, b = f64[0].clone().requires_grad_(True), f64[1].clone().requires_grad_(True)
    feature_distance(a, b).sum().backward()
    a, b = f512[0].clone().requires_grad_(True), f512[1].clone().requires_grad_(True)
    feature_distance(a, b).sum().backward()
    r = raw64.clone().requires_grad_(True)
    y, pooled = bias_relu_pool(r, b64)
    (y.float().square().mean() + pooled.float().square().mean()).backward()


def sec_downsample():
    x = img256.clone().requires_grad_(True)
    S.bilinear_downsample(x, 2, kh, kv).square().mean().backward()


def sec_optim():
    for p in params:
        p.grad = p.detach() * 0.01
    opt.step()
    fr = flow_res.clone().requires_grad_(True)
    total_variation_loss(fr).backward()


def sec_points():
    splat2d(blank, pts, vals, sig, False)
    splat2d_lookup(blank, grid512, query, vals, sig, 512, 512, False)
    nn_argmin(grid512[:, ::4, ::4].contiguous(), query[:, :20000].contiguous())


def sec_small():
    op.upfirdn2d(img128, k4, up=2, pad=(2, 1))
    op.upfirdn2d(img256, k4, down=2, pad=(1, 1))
    style_path.all_demod(w_mod, st_mod, [0.02, 0.02])


for it in range(2):
    for sec in (sec_sampler, sec_perceptual, sec_downsample, sec_optim, sec_points, sec_small):
        try:
            sec()
        except Exception as exc:   # one broken section must not cost the whole capture
            print("prof_misc: %s failed: %r" % (sec.__name__, exc), file=sys.stderr)
torch.cuda.synchronize()
print("prof_misc done")
