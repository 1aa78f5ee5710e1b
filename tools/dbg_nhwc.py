"""Isolate the stages of the NHWC fused-tail backward on one shape (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gangealing_b200.op.upfirdn2d import upfirdn2d_raw
from gangealing_b200.op.fused_act import bias_act_backward_raw
from gangealing_b200.op.modconv import channel_scale_raw
CL = torch.channels_last
torch.manual_seed(0)
def report(name, a, e):
    d = (a.float() - e.float()).abs()
    bad = (d > 1e-3 * e.abs().max()).nonzero()
    print("%-28s max err %.3e (ref %.3e) bad %d first %s" % (name, d.max().item(), e.abs().max().item(), bad.shape[0],
          bad[:4].tolist() if bad.shape[0] else ""), flush=True)
    if bad.shape[0]:
        for dim in range(4):
            u = torch.unique(bad[:, dim])
            print("    dim%d: %d distinct, min %d max %d" % (dim, u.numel(), u.min().item(), u.max().item()))
for (n, c, h, w) in [(2, 128, 256, 256)]:
    print("shape", (n, c, h, w))
    k = torch.tensor([1., 3., 3., 1.]); k = (k[:, None] * k[None, :]); k = (k / k.sum() * 4).cuda()
    g = torch.randn(n, c, h, w, device="cuda").contiguous(memory_format=CL)
    out = torch.randn(n, c, h, w, device="cuda").contiguous(memory_format=CL)
    gx, gb = bias_act_backward_raw(g, out, 0.2, 2 ** 0.5, True)
    e = torch.where(out > 0, g, g * 0.2) * 2 ** 0.5
    report("bias_act_bwd gx", gx, e)
    report("bias_act_bwd gbias", gb.reshape(1, c, 1, 1), e.sum((0, 2, 3)).reshape(1, c, 1, 1))
    for pad in [(2, 2, 2, 2), (1, 1, 1, 1)]:
        y = upfirdn2d_raw(g, k, (1, 1), (1, 1), pad)
        e = F.conv2d(F.pad(g.contiguous(), pad).reshape(n * c, 1, h + pad[2] + pad[3], w + pad[0] + pad[1]),
                     torch.flip(k, [0, 1])[None, None]).reshape(n, c, y.shape[2], y.shape[3])
        report("blur pad %s" % (pad,), y, e)
        for _ in range(3):
            y2 = upfirdn2d_raw(g, k, (1, 1), (1, 1), pad)
            if not torch.equal(y, y2):
                print("    NONDETERMINISTIC rerun differs: %d elems" % (y != y2).sum().item())
    s = torch.rand(n, c, device="cuda") + 0.5
    yy = torch.randn(n, c, h, w, device="cuda").contiguous(memory_format=CL)
    o, dot = channel_scale_raw(g, s, y=yy)
    report("channel_scale out", o, g * s[:, :, None, None])
    report("channel_scale dot", dot.reshape(n, c, 1, 1), (g * yy).sum((2, 3)).reshape(n, c, 1, 1))

print("---- full fused tail, NHWC vs NCHW paths")
torch.backends.cudnn.allow_tf32 = False
from gangealing_b200 import op
for (n, c, h, w) in [(2, 128, 257, 257), (2, 32, 257, 257), (1, 128, 257, 257), (2, 128, 129, 129), (2, 128, 257, 65), (2, 128, 65, 257)]:
    print("shape", (n, c, h, w))
    k = torch.tensor([1., 3., 3., 1.]); k = (k[:, None] * k[None, :]); k = (k / k.sum() * 4).cuda()
    x = torch.randn(n, c, h, w, device="cuda")
    noise = torch.randn(n, 1, h - 1, w - 1, device="cuda")
    nw = torch.randn(1, device="cuda"); b = torch.randn(c, device="cuda"); rs = torch.rand(n, c, device="cuda") + 0.5
    go = torch.randn(n, c, h - 1, w - 1, device="cuda")
    res = []
    for cl in (False, True):
        xx = (x.contiguous(memory_format=CL) if cl else x.clone()).requires_grad_(True)
        rr = rs.clone().requires_grad_(True)
        y = op.blur_noise_bias_act(xx, k, (1, 1), noise, nw, b, row_scale=rr)
        gx, grs = torch.autograd.grad(y, [xx, rr], go.contiguous(memory_format=CL) if cl else go)
        res.append((y, gx, grs))
    report("fwd", res[1][0], res[0][0])
    report("grad x", res[1][1], res[0][1])
    report("grad rs", res[1][2].reshape(n, c, 1, 1), res[0][2].reshape(n, c, 1, 1))
    g = torch.randn(n, c, h, w, device="cuda").contiguous(memory_format=CL)
    yy = torch.randn(n, c, h, w, device="cuda").contiguous(memory_format=CL)
    o, dot = channel_scale_raw(g, rs, y=yy)
    report("channel_scale out", o, g * rs[:, :, None, None])
    report("channel_scale dot", dot.reshape(n, c, 1, 1), (g * yy).sum((2, 3)).reshape(n, c, 1, 1))
