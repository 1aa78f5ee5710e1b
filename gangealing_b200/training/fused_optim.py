"""FusedAdamEMA -- Adam for every trainable tensor of the step (STN + latent learner) and the EMA of the STN as ONE
multi-tensor kernel (csrc/optim.cu `gg_adam_ema_step`), SURVEY.md 8(f) rank 3.

reference: train.py:126-134 (`t_optim.step()`, `ll_optim.step()`, `accumulate(t_ema, t_module)`), optimisers built at
train.py:204-205 (`optim.Adam(..., betas=(0.9, 0.999), eps=1e-8)`), EMA models/__init__.py:19-24.

A torch.optim.Optimizer subclass: `param_groups` (one learning rate each -- a device scalar, so a captured CUDA graph follows
the schedule) and a per-parameter `state` with torch.optim.Adam's keys (`step`, `exp_avg`, `exp_avg_sq`), hence
`state_dict()` / `load_state_dict()` exchange checkpoints with the reference's optimisers.
"""
import torch

from .. import _lib

_CHUNK = 65536


class FusedAdamEMA(torch.optim.Optimizer):
    def __init__(self, param_groups, betas=(0.9, 0.999), eps=1e-8, ema_pairs=None, ema_decay=0.999):
        """param_groups: [{"params": [...], "lr": float}, ...]; ema_pairs: {trainable parameter: its EMA twin}."""
        defaults = dict(lr=1e-3, betas=betas, eps=eps)
        super().__init__(param_groups, defaults)
        self.ema = dict(ema_pairs or {})
        self.ema_decay = float(ema_decay)
        params = [p for g in self.param_groups for p in g["params"]]
        if not params:
            raise ValueError("FusedAdamEMA needs parameters")
        self.device = params[0].device
        _lib.require_cuda(*params)
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() and not p.is_contiguous(memory_format=torch.channels_last):
                raise RuntimeError("FusedAdamEMA: parameters must be dense fp32 tensors")
        # one device scalar per group holds the learning rate; state[0] of `_state3` is the shared step counter
        self._lr = [torch.tensor(float(g["lr"]), device=self.device) for g in self.param_groups]
        self._state3 = torch.zeros(3, device=self.device)
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state[p]
                st["step"] = self._state3[0:1].view(())          # a view: every parameter shares the counter
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        # CTA -> (tensor, chunk) maps are static (sizes never change)
        bt, bc = [], []
        self._order = []
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                ti = len(self._order)
                self._order.append((p, gi))
                for c in range((p.numel() + _CHUNK - 1) // _CHUNK):
                    bt.append(ti)
                    bc.append(c)
        self._blocks = len(bt)
        self._block_tensor = torch.tensor(bt, dtype=torch.int32, device=self.device)
        self._block_chunk = torch.tensor(bc, dtype=torch.int32, device=self.device)
        self._table_host = torch.zeros((len(self._order), 7), dtype=torch.int64).pin_memory()
        self._table = torch.zeros((len(self._order), 7), dtype=torch.int64, device=self.device)
        self._table_key = None
        self._ship = {"host": self._table_host, "dev": self._table}

    # ---- schedule ------------------------------------------------------------------------------------------------
    def set_lr(self, group, value):
        """value: float or 0-dim tensor; lands in the device scalar the kernel reads (graph-replay safe)."""
        if torch.is_tensor(value):
            self._lr[group].copy_(value, non_blocking=True)
        else:
            self._lr[group].fill_(float(value))
        self.param_groups[group]["lr"] = value if not torch.is_tensor(value) else float("nan")

    def lr_tensor(self, group):
        return self._lr[group]

    # ---- checkpoints ---------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """torch's loader REPLACES every state tensor by a copy of the checkpoint's and hands each parameter its own `step`;
        the kernel's pointer table (and a captured CUDA graph) holds the addresses of the moment tensors created in
        `__init__`, and the step count lives in one shared device scalar.  So: load, then copy the loaded values INTO the
        original tensors (any layout / dtype / device of the checkpoint: the reference's optimisers keep NCHW-strided fp32
        moments) and put those back -- no address changes, nothing to re-ship, graph replays stay valid."""
        own = {p: (st["exp_avg"], st["exp_avg_sq"]) for p, st in self.state.items() if "exp_avg" in st}
        super().load_state_dict(state_dict)
        steps = []
        for p, (m, v) in own.items():
            st = self.state[p]
            if "exp_avg" in st:
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                steps.append(float(st.get("step", 0.0)))
            else:                      # a checkpoint taken before the first step has no state for this tensor
                m.zero_()
                v.zero_()
                steps.append(0.0)
            st["exp_avg"], st["exp_avg_sq"] = m, v
            st["step"] = self._state3[0:1].view(())
        if steps and min(steps) != max(steps):
            raise RuntimeError("FusedAdamEMA.load_state_dict: the checkpoint's tensors disagree on the step count (%g .. %g); "
                               "the fused step keeps ONE counter for all of them" % (min(steps), max(steps)))
        self._state3[0:1].fill_(steps[0] if steps else 0.0)
        for i, g in enumerate(self.param_groups):   # the loaded learning rates go where the kernel reads them
            lr = g["lr"]
            if torch.is_tensor(lr):
                self._lr[i].copy_(lr)
            elif lr == lr:             # nan marks "set from a device tensor" (set_lr): keep the scalar as it is
                self._lr[i].fill_(float(lr))

    # ---- step ----------------------------------------------------------------------------------------------------
    def _refresh_table(self):
        """Pointer table of this step's tensors.  Gradient tensors are re-created by autograd every iteration (fixed
        addresses inside a captured graph's pool, DDP bucket views otherwise), so the table is rebuilt whenever a pointer
        moved and shipped with one small pinned-memory copy."""
        rows, key = [], []
        for p, gi in self._order:
            if p.grad is None:
                raise RuntimeError("FusedAdamEMA.step: a parameter has no gradient (the fused step updates every tensor)")
            g = p.grad
            if g.dtype != torch.float32 or g.stride() != p.stride():
                g = g.to(torch.float32).contiguous(memory_format=torch.preserve_format) if g.stride() == p.stride() else \
                    torch.empty_like(p).copy_(g)
                p.grad = g
            st = self.state[p]
            e = self.ema.get(p)
            if e is not None and e.stride() != p.stride():
                raise RuntimeError("FusedAdamEMA: an EMA twin is laid out differently from its parameter")
            rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         0 if e is None else e.data_ptr(), p.numel(), self._lr[gi].data_ptr()))
            key.append(g.data_ptr())
        key = tuple(key)
        if key != self._table_key:
            _lib.ship_table(self._ship, torch.tensor(rows, dtype=torch.int64))
            self._table_key = key

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("FusedAdamEMA does not take a closure")
        self._refresh_table()
        g0 = self.param_groups[0]
        rc = _lib.load().gg_adam_ema_step(self._table.data_ptr(), self._block_tensor.data_ptr(), self._block_chunk.data_ptr(),
                                          self._blocks, _CHUNK, self._state3.data_ptr(), g0["betas"][0], g0["betas"][1], g0["eps"],
                                          self.ema_decay, _lib.stream())
        _lib.check(rc, "gg_adam_ema_step")


def split_adam_state_dict(state_dict):
    """One optimiser `state_dict` with G parameter groups -> G single-group dicts in torch.optim.Adam's layout, parameter
    indices renumbered from 0: what the reference's checkpoints hold as `t_optim` and `ll_optim` (train.py:20-27)."""
    out, state = [], state_dict["state"]
    for group in state_dict["param_groups"]:
        ids = list(group["params"])
        g = {k: v for k, v in group.items() if k != "params"}
        g["params"] = list(range(len(ids)))
        out.append({"state": {new: state[old] for new, old in enumerate(ids) if old in state}, "param_groups": [g]})
    return out


def merge_adam_state_dicts(state_dicts):
    """Inverse of `split_adam_state_dict`: the reference's per-network optimiser dicts (each may itself hold several groups)
    -> one dict whose groups follow each other, parameter indices renumbered consecutively."""
    state, groups, base = {}, [], 0
    for sd in state_dicts:
        count = 0
        for group in sd["param_groups"]:
            g = {k: v for k, v in group.items() if k != "params"}
            g["params"] = [base + i for i in group["params"]]
            count = max([count] + [i + 1 for i in group["params"]])
            groups.append(g)
        for i, st in sd["state"].items():
            state[base + int(i)] = st
        base += count
    return {"state": state, "param_groups": groups}
