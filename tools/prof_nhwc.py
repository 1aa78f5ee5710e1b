"""Tiny driver for ncu: the channels-last fused kernels at the 256^2 / 128^2 layer shapes of the bench (per-GPU batch B),
storage type DT (f32 | bf16).  Each kernel family runs twice (the first launch warms caches / the tensor-map cache)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200.op import nhwc
from gangealing_b200 import _lib
from gangealing_b200.op.upfirdn2d import grad_pad
B = int(os.environ.get("B", "32"))
dt = torch.bfloat16 if os.environ.get("DT", "f32") == "bf16" else torch.float32
dev = "cuda"
CL = torch.channels_last
k4 = torch.tensor([1., 3., 3., 1.]); k4 = k4[None] * k4[:, None]; k4 = (k4 / k4.sum() * 4).to(dev)
C, H = 128, 256
raw_up = torch.randn(B, C, H + 1, H + 1, device=dev).to(dt).contiguous(memory_format=CL)
raw = torch.randn(B, C, H, H, device=dev).to(dt).contiguous(memory_format=CL)
g = torch.randn(B, C, H, H, device=dev).to(dt).contiguous(memory_format=CL)
noise = torch.randn(B, 1, H, H, device=dev)
nw = torch.tensor([0.1], device=dev); bias = torch.randn(C, device=dev)
demod = torch.rand(B, C, device=dev) + 0.5; s_next = torch.randn(B, C, device=dev) + 1.0
wm = torch.randn(B, 3, C, device=dev) / C ** 0.5; rgb_bias = torch.randn(3, device=dev); skip = torch.randn(B, 3, H, H, device=dev)
g_rgb = torch.randn(B, 3, H, H, device=dev)
gp = grad_pad(H + 1, H + 1, H, H, 4, 4, (1, 1), (1, 1), (1, 1, 1, 1))
for _ in range(2):
    # forward, no backward pending (generator pass 1): blur tail emits only the next conv's input; last layer only the image
    nhwc.blur(raw_up, k4, (1, 1, 1, 1), mode=1, noise=noise, noise_weight=nw, bias=bias, row_scale=demod, scale2=s_next,
              want_out=False, want_out2=True, negative_slope=0.2, gain=2 ** 0.5)
    nhwc.styled_tail(raw, noise, nw, bias, demod, None, wm, rgb_bias, skip, False, 0.2, 2 ** 0.5)
    # forward with a backward pending (pass 2): dual emit
    out, xs, _ = nhwc.blur(raw_up, k4, (1, 1, 1, 1), mode=1, noise=noise, noise_weight=nw, bias=bias, row_scale=demod, scale2=s_next,
                           want_out=True, want_out2=True, negative_slope=0.2, gain=2 ** 0.5)
    o2, xs2, rgb = nhwc.styled_tail(raw, noise, nw, bias, demod, s_next, wm, rgb_bias, skip, True, 0.2, 2 ** 0.5)
    # backward
    nhwc.styled_tail_backward(g, g_rgb, o2, raw, s_next, demod, wm, True, True, True, 0.2, 2 ** 0.5)
    g_t = nhwc.styled_tail_backward(g, None, out, None, s_next, None, None, True, False, False, 0.2, 2 ** 0.5)[0]
    nhwc.blur(g_t, _lib.flipped_filter(k4), gp, mode=2, row_scale=demod, mul=raw_up, want_dot=True)
    # STN trunk family
    nhwc.blur(raw_up, k4, (1, 1, 1, 1), mode=0)
    y = nhwc.noise_bias_act(raw, None, None, bias, None, 0.2, 2 ** 0.5)
    nhwc.bias_act_backward(g, y, 0.2, 2 ** 0.5, True)
torch.cuda.synchronize()
