#!/usr/bin/env python
"""Measured lines for BASELINE configs 4 and 5 (the two configurations bench.py does not time), one JSON line each.

  python tools/configbench.py --config 5 [--batch 5] [--dtype f32|bf16]                 # 1 GPU
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/configbench.py --config 4|5 ...   # N GPUs

config 5  LSUN Cars K=4 clustering (reference scripts/training/lsun_cars.sh:5-7, models/losses/loss.py:78-92): one training
          iteration of `gangealing_cluster_loss` -- per rank B latents -> G x2 (B, then B*K), STN on 2B inputs (flips) -> 2BK
          warps sampled from the full-resolution image, perceptual loss on 2BK pairs, min over 2K, backward, Adam x2, EMA --
          as a whole-step CUDA graph.  DDP over N ranks (204 MB of STN gradients).  -> training images/s (B per rank per step).
config 4  CelebA-HQ 512^2 propagation (applications/propagate_to_images.py:44-78): per frame the similarity+flow STN at
          supersize 512 / output_resolution 512, `uncongeal_points` of P key points (a disc rendered at R = 256 / 1024:
          P = 25 233 / 403 533) and `splat_points` at sigma 0.3 / 1.3 (object + mask splats, alpha blend).  Frames are sharded
          across ranks (no data-path collective; the reference all_gathers the finished frames for the video writer,
          mixed_reality.py:28-33 -- that gather of finished frames is included).  -> frames/s and, per stage, device ms.
Timing: CUDA events, max over ranks, inputs resident on the device, >= 3 warm-up iterations; a step's activations exceed L2.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        from gangealing_b200.training import distributed as gdist
        gdist.setup_distributed("nccl")
    return world, int(os.environ.get("RANK", "0")), "cuda:%d" % local


def sync(world):
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(vals, dev, world):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def config5(args, world, rank, dev):
    from gangealing_b200.training import TrainConfig, Trainer
    torch.backends.cudnn.benchmark = True
    cfg = TrainConfig(batch=args.batch, num_heads=4, flips=True, ndirs=5, inject=6, sample_from_full_res=True,
                      padding_mode="reflection", dtype=args.dtype,
                      grad_compression="bf16" if args.dtype == "bf16" else "none")
    tr = Trainer(cfg, dev, distributed=world > 1)
    for _ in range(max(args.warmup, 3)):
        tr.step()
    sync(world)
    tr.capture(warmup=2)
    for _ in range(2):
        tr.step()
    sync(world)
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(args.steps):
        out = tr.step()
    en.record()
    sync(world)
    (ms,) = max_over_ranks([st.elapsed_time(en)], dev, world)
    line = {"metric": "gangealing_cluster_train_images_per_sec_256", "value": args.batch * world * args.steps / (ms / 1e3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "dtype": args.dtype,
            "scaling": "weak", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: LSUN Cars K=4 clustering step (gangealing_cluster_loss: num_heads 4, flips, "
                                   "ndirs 5, inject 6, sample_from_full_res, reflection padding), StyleGAN2-256 + similarity+flow STN "
                                   "@128, perceptual VGG16 loss on 2*B*K pairs, Adam x2, EMA; synthetic latents, seeded random weights",
                       "per_gpu_batch": args.batch, "warps_per_step_per_gpu": 2 * 4 * args.batch,
                       "step_mode": "whole-step CUDA graph replay", "parallelism": "dp%d" % world},
            "losses": {k: float(v.detach()) for k, v in out.items()}}
    return line, tr


def config4(args, world, rank, dev):
    import torch.distributed as dist
    from gangealing_b200.stn import get_stn
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    stn = get_stn(["similarity", "flow"], flow_size=128, supersize=512, channel_multiplier=0.5, num_heads=1).to(dev).eval()
    stn.to(memory_format=torch.channels_last)
    for m in stn.modules():
        if hasattr(m, "channels_last") and hasattr(m, "stn_in_size"):
            m.channels_last = True
    with torch.no_grad():   # a non-identity warp (the heads are zero-initialised: warping_heads.py:28-30,164-165)
        g = torch.Generator().manual_seed(1)
        for name, prm in stn.named_parameters():
            if "warp_head" in name:
                prm.copy_((0.05 * torch.randn(prm.shape, generator=g)).to(dev))
    B = args.batch
    gen = torch.Generator(device=dev).manual_seed(10 + rank)          # each rank owns disjoint frames
    frames = (torch.rand(B, 3, 512, 512, device=dev, generator=gen) * 2 - 1).clamp(-1, 1)
    rows = []
    for R in (256, 1024):
        ys, xs = torch.meshgrid(torch.arange(float(R)), torch.arange(float(R)), indexing="ij")
        disc = ((ys - R / 2) ** 2 + (xs - R / 2) ** 2) < (0.35 * R) ** 2
        pts = (torch.stack([xs[disc], ys[disc]], dim=1)[None] * (127.0 / (R - 1))).repeat(B, 1, 1).to(dev).contiguous()
        P = pts.shape[1]
        colors = torch.randn(B, P, 3, device=dev)
        for sigma in (0.3, 1.3):
            def local_work():
                with torch.no_grad():
                    return stn.uncongeal_and_splat(frames, pts, colors, sigma, 0.75, output_resolution=512,
                                                   normalize_input_points=True, padding_mode="border")[0]

            def stn_only():
                with torch.no_grad():
                    return stn(frames, return_warp=True, output_resolution=512, padding_mode="border")[0]

            # the per-batch work as CUDA graphs (device time: ~150 launches per batch would otherwise be launch-bound)
            for _ in range(max(args.warmup, 3)):
                local_work()
                stn_only()
            sync(world)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g_all, g_stn = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_all, stream=side):
                img = local_work()
            with torch.cuda.graph(g_stn, stream=side):
                warped = stn_only()
            torch.cuda.current_stream().wait_stream(side)
            gathered = [torch.empty_like(img) for _ in range(world)] if world > 1 else None

            def frame_batch():
                g_all.replay()
                if world > 1:
                    dist.all_gather(gathered, img)                     # finished frames, as mixed_reality.py:28-33

            for _ in range(3):
                frame_batch()
                g_stn.replay()
            sync(world)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(args.steps):
                frame_batch()
            ev[1].record()
            for _ in range(args.steps):
                g_stn.replay()
            ev[2].record()
            sync(world)
            ms_all, ms_stn = max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])], dev, world)
            per = ms_all / args.steps
            rows.append({"points": P, "sigma": sigma, "frames_per_s": B * world * args.steps / (ms_all / 1e3),
                         "ms_per_batch": per, "stn_ms_per_batch": ms_stn / args.steps,
                         "lookup_splat_blend_ms_per_batch": per - ms_stn / args.steps})
            del g_all, g_stn
    best = max(r["frames_per_s"] for r in rows)
    line = {"metric": "gangealing_propagate_frames_per_sec_512", "value": best, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "dtype": "f32", "scaling": "weak", "data": "synthetic",
            "config": {"workload": "BASELINE config 4: CelebA-HQ 512^2 propagation (flow STN supersize 512 -> uncongeal_points -> "
                                   "splat_points, alpha blend), frames sharded across ranks + all_gather of finished frames",
                       "frames_per_gpu_per_batch": B, "parallelism": "frames sharded x%d" % world,
                       "step_mode": "per-batch work replayed from a CUDA graph; the all_gather of finished frames eager"},
            "rows": rows}
    return line, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--batch", type=int, default=None, help="config 5: latents per GPU (reference 5); config 4: frames per GPU per batch")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 5 if args.config == 5 else 8
    world, rank, dev = setup()
    line, tr = (config5 if args.config == 5 else config4)(args, world, rank, dev)
    if rank == 0:
        print(json.dumps(line), flush=True)
        if args.json:
            os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
            with open(args.json, "w") as f:
                json.dump(line, f, indent=1)
    if world > 1:
        import threading
        import torch.distributed as dist
        wd = threading.Timer(20.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.barrier()
        if tr is not None:
            tr.release_graph()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
