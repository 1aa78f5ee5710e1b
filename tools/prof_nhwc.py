"""Tiny driver for ncu: the channels-last kernels at the 256^2 layer shape of the bench (per-GPU batch B)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200 import op
from gangealing_b200.op.modconv import channel_scale_raw, _ToRGB
from gangealing_b200.op.fused_act import bias_act_backward_raw
B = int(os.environ.get("B", "32"))
dev = "cuda"
CL = torch.channels_last
k4 = torch.tensor([1., 3., 3., 1.]); k4 = k4[None] * k4[:, None]; k4 = (k4 / k4.sum() * 4).to(dev)
x = torch.randn(B, 128, 257, 257, device=dev).contiguous(memory_format=CL)
y = torch.randn(B, 128, 256, 256, device=dev).contiguous(memory_format=CL)
y2 = torch.randn(B, 128, 256, 256, device=dev).contiguous(memory_format=CL)
noise = torch.randn(B, 1, 256, 256, device=dev)
nw = torch.tensor([0.1], device=dev); bias = torch.randn(128, device=dev); rs = torch.rand(B, 128, device=dev) + 0.5
wm = torch.randn(B, 3, 128, device=dev); b3 = torch.randn(1, 3, 1, 1, device=dev); skip = torch.randn(B, 3, 256, 256, device=dev)
for _ in range(2):
    op.upfirdn2d(x, k4, pad=(1, 1))
    op.blur_noise_bias_act(x, k4, (1, 1), noise, nw, bias, row_scale=rs)
    op.noise_bias_act(y, noise, nw, bias, row_scale=rs)
    channel_scale_raw(y, rs, y=y2)
    bias_act_backward_raw(y, y2, 0.2, 1.4, True)
    yy = y.clone().requires_grad_(True); w2 = wm.clone().requires_grad_(True)
    o = _ToRGB.apply(yy, w2, b3, skip)
    o.backward(torch.ones_like(o))
torch.cuda.synchronize()
