"""CPU, world_size 2 over gloo: the data-parallel logic of the train step (per-rank latents, DDP gradient
all-reduce keeps replicas identical, rank-0 loss reduce)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from oracle import opset
    from gangealing_b200.training import TrainConfig, Trainer
    from gangealing_b200.training import distributed as gdist
    assert gdist.setup_distributed("gloo")
    cfg = TrainConfig(gen_size=64, flow_size=64, dim_latent=16, n_mlp=1, batch=1, inject=3, seed=3)
    tr = Trainer(cfg, "cpu", ops=opset.cpu_ops(), distributed=True)
    z_probe = torch.randn(2)                      # the per-rank RNG stream differs (train.py:193)
    for _ in range(2):
        out = tr.step()
    flat = torch.cat([p.detach().reshape(-1) for p in tr.t_module.parameters()])
    gathered = gdist.all_gather(flat[None], cat=True)
    probes = gdist.all_gather(z_probe[None], cat=True)
    if rank == 0:
        ret["replicas_equal"] = bool(torch.equal(gathered[0], gathered[1]))
        ret["latents_differ"] = bool(not torch.equal(probes[0], probes[1]))
        ret["loss_keys"] = sorted(out.keys())
        ret["loss_finite"] = bool(all(torch.isfinite(v) for v in out.values()))
    gdist.synchronize()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_data_parallel_step_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = 29500 + (os.getpid() % 2000)
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(560)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert ret["replicas_equal"], "DDP replicas diverged"
        assert ret["latents_differ"]
        assert ret["loss_keys"] == ["f", "p", "tv"] and ret["loss_finite"]
