"""One cluster-classifier training iteration (reference train_cluster_classifier.py:79-107), the second half of BASELINE
config 5: fake images are assigned to clusters by the frozen clustering STN (`assign_fake_images_to_clusters`, no gradient),
and the classifier learns to predict (cluster, flip) from the unaligned image with a cross-entropy loss.

Everything device-side is the hot path's own kernels: the generator and the STN as in `Trainer`, the classifier trunk = the
similarity STN's trunk (cluster_classifier.py), Adam through `gg_adam_ema_step` (no EMA twin here).  No host sync in `step()`.
"""
import torch
from torch import nn, optim

from ..cluster_classifier import ResnetClassifier, accuracy
from . import distributed as gdist
from .losses import assign_fake_images_to_clusters
from .schedule import decaying_cosine_lr


class ClassifierTrainer:
    """Built on a `Trainer` (which owns the frozen generator, the EMA STN, the latent learner, the perceptual loss and the
    G->STN resize): `ClassifierTrainer(trainer).step()` runs one iteration and returns the reference's loss dict
    (`cross_entropy`, `acc@1`, `acc@2`, `head_i`, `pred_head_i`) as device scalars."""

    def __init__(self, trainer, cls_lr=1e-3, real_size=None, init_from_stn=True, distributed=False, ops=None):
        """`ops`: the op set the `trainer` was built with (None = the sm_100a kernels; tests pass the oracle's CPU set)."""
        cfg = trainer.cfg
        self.trainer, self.cfg, self.device = trainer, cfg, trainer.device
        self.total_clusters = cfg.num_heads * (1 + int(cfg.flips))
        native = ops is None
        # reference :169-170: ResnetClassifier(flow_size, stn_channel_multiplier, num_heads * (1 + flips), supersize=real_size)
        self.classifier = ResnetClassifier(cfg.flow_size, channel_multiplier=cfg.stn_channel_multiplier,
                                           num_heads=self.total_clusters, supersize=real_size or cfg.flow_size,
                                           ops=ops).to(self.device)
        if init_from_stn:   # reference :189-193: start from the similarity STN's trunk
            first = trainer.t_ema.stns[0] if hasattr(trainer.t_ema, "stns") else trainer.t_ema
            self.classifier.load_state_dict(first.state_dict(), strict=False)
        if trainer.generator.channels_last:
            self.classifier.to(memory_format=torch.channels_last)
            self.classifier.channels_last = True
            self.classifier.act_dtype = trainer.generator.act_dtype
        for p in self.classifier.parameters():
            p.requires_grad = True
        self.module = self.classifier
        if distributed:
            ids = [torch.cuda.current_device()] if torch.device(self.device).type == "cuda" else None
            self.classifier = nn.parallel.DistributedDataParallel(self.classifier, device_ids=ids, broadcast_buffers=False)
        self.cls_lr = float(cls_lr)
        self.xent = nn.CrossEntropyLoss()
        on_gpu = torch.device(self.device).type == "cuda"
        if on_gpu and native and cfg.fused_optimizer:
            from .fused_optim import FusedAdamEMA
            self.optim = FusedAdamEMA([{"params": list(self.module.parameters()), "lr": self.cls_lr}], betas=(0.9, 0.999), eps=1e-8)
            self.lr_t = self.optim.lr_tensor(0)
        else:
            self.lr_t = torch.tensor(self.cls_lr, device=self.device) if on_gpu else None
            self.optim = optim.Adam(self.module.parameters(), lr=self.lr_t if on_gpu else self.cls_lr, betas=(0.9, 0.999), eps=1e-8,
                                    fused=on_gpu, capturable=on_gpu)
        self.psi = torch.tensor(0.0, device=self.device)     # reference :60: the truncation is fully annealed by now

    def set_iteration(self, i, period=37500, tm=2, decay=0.9):
        """Learning rate of iteration i: `cls_sched.step(i / period)` of DecayingCosineAnnealingWarmRestarts (reference :106)."""
        lr = decaying_cosine_lr(i / period, self.cls_lr, tm, decay)
        if self.lr_t is not None:
            self.lr_t.fill_(lr)
        else:
            for group in self.optim.param_groups:
                group["lr"] = lr
        return lr

    def checkpoint(self, iteration=None):
        """The reference's classifier checkpoint (train_cluster_classifier.py:25-29): `classifier`, `g_ema`, `t_ema`, `ll`,
        `cls_optim` (torch.optim.Adam's layout; the fused optimiser holds one group here, so its dict already has it)."""
        t = self.trainer
        return {"classifier": self.module.state_dict(), "g_ema": t.generator.state_dict(), "t_ema": t.t_ema.state_dict(),
                "ll": t.ll_module.state_dict(), "cls_optim": self.optim.state_dict(), "iteration": iteration}

    def load_checkpoint(self, ckpt):
        """Restore what train_cluster_classifier.py:180-204 restores: generator, clustering STN and latent learner always; the
        classifier and its optimiser when the checkpoint holds them (-> True), else the classifier keeps its initialisation
        from the similarity STN (-> False)."""
        t = self.trainer
        if getattr(t, "_graph", None) is not None:    # as Trainer.load_checkpoint: derived generator weights are capture constants
            t.release_graph()
        t.generator.load_state_dict(ckpt["g_ema"])
        t.t_ema.load_state_dict(ckpt["t_ema"])
        t.ll_module.load_state_dict(ckpt["ll"])
        if "classifier" not in ckpt:
            first = t.t_ema.stns[0] if hasattr(t.t_ema, "stns") else t.t_ema
            self.module.load_state_dict(first.state_dict(), strict=False)
            return False
        self.module.load_state_dict(ckpt["classifier"])
        self.optim.load_state_dict(ckpt["cls_optim"])
        if self.lr_t is not None and not hasattr(self.optim, "lr_tensor"):
            # capturable torch Adam: the learning rate must remain THE device scalar (the fused optimiser's loader does this itself)
            for group in self.optim.param_groups:
                self.lr_t.copy_(group["lr"]) if torch.is_tensor(group["lr"]) else self.lr_t.fill_(float(group["lr"]))
                group["lr"] = self.lr_t
        if ckpt.get("iteration") is not None:
            self.set_iteration(int(ckpt["iteration"]))
        return True

    def losses(self, z=None):
        t, cfg = self.trainer, self.cfg
        with torch.no_grad():   # image formation and cluster assignment are not differentiated (reference :84-89)
            assigned, _, _, _, resized, distance = assign_fake_images_to_clusters(
                t.generator, t.t_ema, t.ll_module, t.loss_fn, t.resize_fake2stn, self.psi, cfg.batch, cfg.dim_latent, True,
                cfg.num_heads, cfg.flips, self.device, sample_from_full_res=cfg.sample_from_full_res, z=z,
                padding_mode=cfg.padding_mode)
        logits = self.classifier(resized[:cfg.batch])
        out = {"cross_entropy": self.xent(logits, assigned.indices),
               "acc@1": accuracy(logits, -distance), "acc@2": accuracy(logits, -distance, k=2)}
        with torch.no_grad():   # reference :96-97 uses torch.bincount, which reads its maximum back to the host; same counts here
            heads = torch.arange(self.total_clusters, device=logits.device)
            gt = (assigned.indices[:, None] == heads).sum(0).div(float(cfg.batch))
            pred = (logits.argmax(dim=1)[:, None] == heads).sum(0).div(float(cfg.batch))
        for c in range(self.total_clusters):
            out["head_%d" % c], out["pred_head_%d" % c] = gt[c], pred[c]
        return out

    def step(self, z=None):
        out = self.losses(z)
        self.optim.zero_grad(set_to_none=True)
        out["cross_entropy"].backward()
        self.optim.step()
        return gdist.reduce_loss_dict({k: v.detach() for k, v in out.items()})
