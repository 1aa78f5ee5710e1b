"""Tiny driver for ncu: runs the band kernels (plain blur, fused tail) at the 256^2 layer shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200 import op
B = int(os.environ.get("B", "5"))
dev = "cuda"
k4 = torch.tensor([1., 3., 3., 1.]); k4 = k4[None] * k4[:, None]; k4 = (k4 / k4.sum() * 4).to(dev)
x = torch.randn(B, 128, 257, 257, device=dev)
noise = torch.randn(B, 1, 256, 256, device=dev)
nw = torch.tensor([0.1], device=dev); bias = torch.randn(128, device=dev)
for _ in range(3):
    op.upfirdn2d(x, k4, pad=(1, 1))
    op.blur_noise_bias_act(x, k4, (1, 1), noise, nw, bias)
    y = torch.randn(B, 128, 256, 256, device=dev)
    op.noise_bias_act(y, noise, nw, bias)
torch.cuda.synchronize()
