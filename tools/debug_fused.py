"""Stage-by-stage check of the fused blur tail's backward against the oracle (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200 import _lib
from gangealing_b200.op import nhwc
from gangealing_b200.op.upfirdn2d import grad_pad
from oracle import stylegan2_ops as so
CL = torch.channels_last
dev = "cuda"
for shape in [(2, 64, 17, 17), (2, 128, 33, 41), (2, 128, 33, 33), (2, 64, 33, 41), (1, 128, 17, 41), (1, 32, 40, 50)]:
    n, c, h, w = shape
    g = torch.Generator().manual_seed(1)
    k = so.make_kernel([1, 3, 3, 1]) * 4
    gt = torch.randn(n, c, h - 1, w - 1, generator=g)
    raw = torch.randn(n, c, h, w, generator=g)
    d = torch.rand(n, c, generator=g) + 0.5
    gp = grad_pad(h, w, h - 1, w - 1, 4, 4, (1, 1), (1, 1), (1, 1, 1, 1))
    ref = so.upfirdn2d_ref_full(gt, torch.flip(k, [0, 1]), 1, 1, 1, 1, *gp)
    ref_dot = (ref * raw).sum(dim=(2, 3))
    ref_out = ref * d[:, :, None, None]
    kd = k.to(dev)
    x = gt.to(dev).contiguous(memory_format=CL)
    o0 = nhwc.blur(x, _lib.flipped_filter(kd), gp, mode=0)[0]
    o2, _, dot = nhwc.blur(x, _lib.flipped_filter(kd), gp, mode=2, row_scale=d.to(dev), mul=raw.to(dev).contiguous(memory_format=CL), want_dot=True)
    e0 = (o0.cpu() - ref).abs()
    e2 = (o2.cpu() - ref_out).abs()
    ed = (dot.cpu() - ref_dot).abs()
    print(shape, "gp", gp, "mode0 err %.2e  mode2 err %.2e  dot err %.2e (mag %.1f)" % (e0.max(), e2.max(), ed.max(), ref_dot.abs().max()))
    if e2.max() > 1e-3:
        idx = (e2 > 1e-3).nonzero()
        print("  bad count", idx.shape[0], "first", idx[:5].tolist(), "rows", sorted(set(idx[:, 2].tolist()))[:20], "cols", sorted(set(idx[:, 3].tolist()))[:20])
    # K1: styled_tail_backward without demod / rgb
    out = torch.randn(n, c, h - 1, w - 1, generator=g)
    gxs = torch.randn(n, c, h - 1, w - 1, generator=g)
    s = torch.randn(n, c, generator=g) + 1
    ref_gt = torch.where(out > 0, gxs * s[:, :, None, None], 0.2 * gxs * s[:, :, None, None]) * 2 ** 0.5
    ref_ds = (gxs * out).sum(dim=(2, 3))
    g1, ds, _, _ = nhwc.styled_tail_backward(gxs.to(dev).contiguous(memory_format=CL), None, out.to(dev).contiguous(memory_format=CL), None,
                                            s.to(dev), None, None, True, False, False, 0.2, 2 ** 0.5)
    print("   K1 g_t err %.2e  d_s err %.2e" % ((g1.cpu() - ref_gt).abs().max(), (ds.cpu() - ref_ds).abs().max()))
