"""Grouped (per-sample weights, reference formulation) vs dense (shared weights, modulate activations) convolution
timing at the generator's layer shapes; fwd and fwd+bwd.  Decides how ModulatedConv2d should call cuDNN."""
import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
dev = "cuda"
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for B in (8, 32):
    for (ci, co, h, up) in [(512, 512, 32, False), (512, 512, 32, True), (512, 512, 64, False), (512, 256, 64, True),
                            (256, 256, 128, False), (256, 128, 128, True), (128, 128, 256, False)]:
        x = torch.randn(B, ci, h, h, device=dev, requires_grad=True)
        wg = torch.randn(B * co, ci, 3, 3, device=dev, requires_grad=True)
        wd = torch.randn(co, ci, 3, 3, device=dev, requires_grad=True)
        if up:
            wgt = torch.randn(B * ci, co, 3, 3, device=dev, requires_grad=True)
            wdt = torch.randn(ci, co, 3, 3, device=dev, requires_grad=True)
            fg = lambda: F.conv_transpose2d(x.view(1, B * ci, h, h), wgt, stride=2, groups=B)
            fd = lambda: F.conv_transpose2d(x, wdt, stride=2)
        else:
            fg = lambda: F.conv2d(x.view(1, B * ci, h, h), wg, padding=1, groups=B)
            fd = lambda: F.conv2d(x, wd, padding=1)
        def bw(f):
            def run():
                y = f(); y.backward(torch.ones_like(y))
            return run
        fl = 2 * B * ci * co * 9 * h * h / 1e9
        tg, td, tgb, tdb = t(fg), t(fd), t(bw(fg)), t(bw(fd))
        print("B=%2d %4d->%4d @%3d up=%d  %6.1f GF | fwd grouped %7.3f ms dense %7.3f ms | fwd+bwd grouped %7.3f dense %7.3f"
              % (B, ci, co, h, up, fl, tg, td, tgb, tdb), flush=True)
