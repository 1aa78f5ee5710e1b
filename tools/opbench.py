"""Per-op microbenchmark on one B200: achieved algorithmic GB/s of each hot-path kernel at the
BASELINE config-2 shapes (per-GPU batch B), CUDA-event timed, inputs rotated through a pool larger than L2.
Usage: python tools/opbench.py [--batch 5] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gangealing_b200 import op  # noqa: E402
from gangealing_b200.op.fused_act import bias_act_backward_raw  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timeit(fn, pools, iters=20, warmup=3):
    """fn(i) runs on input set i (rotating) -> ms per call (mean over iters)."""
    for i in range(warmup):
        fn(i % pools)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for i in range(iters):
        fn(i % pools)
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


def nhwc_family(B, dev, k4, timeit, report, pool_count):
    """The channels-last kernels (csrc/nhwc.cu) at the generator's shapes."""
    from gangealing_b200.op.modconv import channel_scale_raw, _ToRGB
    CL = torch.channels_last
    for C, H in [(128, 256), (256, 128), (512, 64), (512, 32), (512, 16), (512, 8)]:
        hin = H + 1
        nbytes = 4 * B * C * (hin * hin + H * H)
        P = pool_count(nbytes)
        xs = [torch.randn(B, C, hin, hin, device=dev).contiguous(memory_format=CL) for _ in range(P)]
        noise = torch.randn(B, 1, H, H, device=dev)
        nw = torch.tensor([0.1], device=dev)
        bias = torch.randn(C, device=dev)
        rs = torch.rand(B, C, device=dev) + 0.5
        ms = timeit(lambda i: op.upfirdn2d(xs[i], k4, pad=(1, 1)), P)
        report("nhwc blur C=%d %d->%d" % (C, hin, H), nbytes, ms)
        ms = timeit(lambda i: op.blur_noise_bias_act(xs[i], k4, (1, 1), noise, nw, bias, row_scale=rs), P)
        report("nhwc fused blur+noise+bias+act C=%d %d->%d" % (C, hin, H), nbytes + 4 * B * H * H, ms)
        del xs
        ys = [torch.randn(B, C, H, H, device=dev).contiguous(memory_format=CL) for _ in range(P)]
        nb2 = 4 * B * C * H * H * 2
        ms = timeit(lambda i: op.noise_bias_act(ys[i], noise, nw, bias, row_scale=rs), P)
        report("nhwc noise+bias+act C=%d %d^2" % (C, H), nb2 + 4 * B * H * H, ms)
        ms = timeit(lambda i: bias_act_backward_raw(ys[i], ys[(i + 1) % P], 0.2, 1.4, True), P)
        report("nhwc bias_act backward(+bias grad) C=%d %d^2" % (C, H), 4 * B * C * H * H * 3, ms)
        ms = timeit(lambda i: channel_scale_raw(ys[i], rs), P)
        report("nhwc channel_scale C=%d %d^2" % (C, H), nb2, ms)
        ms = timeit(lambda i: channel_scale_raw(ys[i], rs, y=ys[(i + 1) % P]), P)
        report("nhwc channel_scale + row_dot C=%d %d^2" % (C, H), 4 * B * C * H * H * 3, ms)
        wm = torch.randn(B, 3, C, device=dev)
        b3 = torch.randn(1, 3, 1, 1, device=dev)
        skip = torch.randn(B, 3, H, H, device=dev)
        ms = timeit(lambda i: _ToRGB.apply(ys[i], wm, b3, skip), P)
        report("nhwc to_rgb forward C=%d %d^2" % (C, H), 4 * B * H * H * (C + 6), ms)
        g3 = torch.randn(B, 3, H, H, device=dev)
        lib = __import__("gangealing_b200._lib", fromlist=["x"])
        gx = torch.empty_like(ys[0]); gw = torch.empty(B, 3, C, device=dev)
        ws = torch.empty(max(1, lib.load().gg_to_rgb_nhwc_workspace(B, C, H * H) // 4), device=dev)
        ms = timeit(lambda i: lib.check(lib.load().gg_to_rgb_nhwc_backward(gx.data_ptr(), gw.data_ptr(), ws.data_ptr(), g3.data_ptr(),
                    ys[i].data_ptr(), wm.data_ptr(), B, C, H * H, lib.stream()), "to_rgb bwd"), P)
        report("nhwc to_rgb backward C=%d %d^2" % (C, H), 4 * B * H * H * (2 * C + 3), ms)
        ms = timeit(lambda i: ys[(i + 1) % P].copy_(ys[i]), P)
        report("  torch copy_ (roofline probe) C=%d %d^2" % (C, H), nb2, ms)
        del ys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"])
    args = ap.parse_args()
    dt = getattr(torch, args.dtype)
    es = torch.empty(0, dtype=dt).element_size()
    B = args.batch
    dev = "cuda"
    peak, src = peak_gbs()
    k4 = torch.tensor([1., 3., 3., 1.])
    k4 = (k4[None] * k4[:, None]); k4 = (k4 / k4.sum() * 4).to(dev)
    rows = []

    def report(name, nbytes, ms):
        gbs = nbytes / ms / 1e6
        rows.append({"op": name, "ms": ms, "GB/s": gbs, "frac": gbs / peak, "MB": nbytes / 1e6})
        print("%-46s %9.3f ms %9.1f MB %8.1f GB/s  %5.1f%% of %s peak" % (name, ms, nbytes / 1e6, gbs, 100 * gbs / peak, src))

    def pool_count(nbytes):
        return max(2, min(8, int(400e6 // max(nbytes, 1)) + 1))

    if args.layout == "nhwc":
        nhwc_family(B, dev, k4, timeit, report, pool_count)
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
            json.dump({"batch": B, "dtype": "float32", "layout": "nhwc", "peak_gbs": peak, "peak_source": src, "rows": rows},
                      open(args.json, "w"), indent=1)
        return
    for C, H in [(128, 256), (256, 128), (512, 64), (512, 32), (512, 16), (512, 8)]:
        hin = H + 1
        nbytes = es * B * C * (hin * hin + H * H)
        P = pool_count(nbytes)
        xs = [torch.randn(B, C, hin, hin, device=dev, dtype=dt) for _ in range(P)]
        ms = timeit(lambda i: op.upfirdn2d(xs[i], k4, pad=(1, 1)), P)
        report("upfirdn2d blur C=%d %d->%d" % (C, hin, H), nbytes, ms)
        noise = torch.randn(B, 1, H, H, device=dev, dtype=dt)
        nw = torch.tensor([0.1], device=dev)
        bias = torch.randn(C, device=dev)
        ms = timeit(lambda i: op.blur_noise_bias_act(xs[i], k4, (1, 1), noise, nw, bias), P)
        report("fused blur+noise+bias+act C=%d %d->%d" % (C, hin, H), nbytes + es * B * H * H, ms)
        ys = [torch.randn(B, C, H, H, device=dev, dtype=dt) for _ in range(P)]
        nb2 = es * B * C * H * H * 2
        ms = timeit(lambda i: op.fused_leaky_relu(ys[i], bias.to(dt)), P)
        report("fused_leaky_relu C=%d %d^2" % (C, H), nb2, ms)
        ms = timeit(lambda i: op.noise_bias_act(ys[i], noise, nw, bias), P)
        report("noise+bias+act C=%d %d^2" % (C, H), nb2 + es * B * H * H, ms)
        ms = timeit(lambda i: bias_act_backward_raw(ys[i], ys[(i + 1) % P], 0.2, 1.4, True), P)
        report("bias_act backward(+bias grad) C=%d %d^2" % (C, H), es * B * C * H * H * 3, ms)
        ms = timeit(lambda i: ys[(i + 1) % P].copy_(ys[i]), P)
        report("  torch copy_ (roofline probe) C=%d %d^2" % (C, H), nb2, ms)
        del xs, ys
    # RGB skip upsample and its backward
    for H in (128, 64):
        x = torch.randn(B, 3, H, H, device=dev, dtype=dt)
        ms = timeit(lambda i: op.upfirdn2d(x, k4, up=2, pad=(2, 1)), 1)
        report("upfirdn2d rgb up2 %d->%d" % (H, 2 * H), es * B * 3 * (H * H + 4 * H * H), ms)
        g = torch.randn(B, 3, 2 * H, 2 * H, device=dev, dtype=dt)
        ms = timeit(lambda i: op.upfirdn2d(g, k4, down=2, pad=(1, 1)), 1)
        report("upfirdn2d rgb-up bwd dn2 %d->%d" % (2 * H, H), es * B * 3 * (H * H + 4 * H * H), ms)
    # STN blurs
    for C, H, pad in [(64, 128, (2, 2)), (64, 128, (1, 1)), (128, 64, (2, 2)), (512, 32, (2, 2))]:
        x = [torch.randn(B, C, H, H, device=dev, dtype=dt) for _ in range(4)]
        ho = H + 2 * pad[0] - 3
        ms = timeit(lambda i: op.upfirdn2d(x[i], k4, pad=pad), 4)
        report("upfirdn2d stn blur C=%d %d pad%s" % (C, H, pad), es * B * C * (H * H + ho * ho), ms)
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump({"batch": B, "dtype": args.dtype, "peak_gbs": peak, "peak_source": src, "rows": rows},
                  open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
