"""How much of the gradient all-reduce hides under the backward pass: torchrun script.
Profiles a few replays of the captured training step with the PyTorch profiler (CUPTI kernel records, no ncu/nsys needed),
and on rank 0 reports: step time, total NCCL kernel time, NCCL time that overlaps other kernels, EXPOSED NCCL time.
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_overlap.py [--dtype bf16] [--batch 32] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200.training import TrainConfig, Trainer  # noqa: E402
from gangealing_b200.training import distributed as gdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--bucket-mb", type=int, default=25)
    ap.add_argument("--compression", default=None)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    gdist.setup_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.backends.cudnn.benchmark = True
    comp = args.compression or ("bf16" if args.dtype == "bf16" else "none")
    tr = Trainer(TrainConfig(batch=args.batch, dtype=args.dtype, grad_compression=comp, bucket_cap_mb=args.bucket_mb), "cuda:%d" % local,
                 distributed=gdist.get_world_size() > 1)
    for _ in range(3):
        tr.step()
    tr.capture(warmup=2)
    for _ in range(3):
        tr.step()
    torch.cuda.synchronize()
    steps = 4
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            tr.step()
        torch.cuda.synchronize()
    if gdist.get_rank() == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
        ker = [(e.time_range.start, e.time_range.end, e.name) for e in ev if "Memcpy" not in e.name and "Memset" not in e.name]
        ker.sort()
        nccl = [(a, b) for a, b, nm in ker if "nccl" in nm.lower()]
        comp_k = [(a, b) for a, b, nm in ker if "nccl" not in nm.lower()]

        def union(iv):
            out = []
            for a, b in sorted(iv):
                if out and a <= out[-1][1]:
                    out[-1][1] = max(out[-1][1], b)
                else:
                    out.append([a, b])
            return out

        def total(iv):
            return sum(b - a for a, b in iv)

        def intersect(x, y):
            i = j = 0
            t = 0.0
            while i < len(x) and j < len(y):
                lo, hi = max(x[i][0], y[j][0]), min(x[i][1], y[j][1])
                if hi > lo:
                    t += hi - lo
                if x[i][1] < y[j][1]:
                    i += 1
                else:
                    j += 1
            return t
        un, uc = union(nccl), union(comp_k)
        span = (ker[-1][1] - ker[0][0]) / steps
        res = {"world": gdist.get_world_size(), "dtype": args.dtype, "per_gpu_batch": args.batch, "bucket_cap_mb": args.bucket_mb,
               "grad_compression": comp, "steps_profiled": steps,
               "step_span_us": span, "nccl_kernels_per_step": len(nccl) / steps, "nccl_busy_us_per_step": total(un) / steps,
               "nccl_overlapped_us_per_step": intersect(un, uc) / steps,
               "nccl_exposed_us_per_step": (total(un) - intersect(un, uc)) / steps,
               "compute_busy_us_per_step": total(uc) / steps,
               "nccl_kernel_names": sorted({nm for _, _, nm in ker if "nccl" in nm.lower()})}
        print(json.dumps(res))
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
            json.dump(res, open(args.json, "w"), indent=1)
    tr.release_graph()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
