"""Equalised-learning-rate weights of a whole network as multi-tensor launches (csrc/optim.cu `gg_scale_cast_multi`).

reference: every `EqualConv2d` / `EqualLinear` forward computes `self.weight * self.scale` (models/stylegan2/networks.py:121-127,
:146-149), and autograd multiplies the weight gradient by the same scalar in backward.  For the Spatial Transformer that is 62
layers: 124 parameter-sized elementwise launches per training step (~3.6 us each -- 4 % of the step at the reference recipe's
per-GPU batch 5), plus a dtype cast each way when the trunk runs on bf16 activations.

`WeightScaler` serves those products from GROUPS of layers: the first layer of a group that runs in a step materialises the
scaled (and cast) weights of the whole group with ONE launch, and the group's backward turns all of its weight gradients into
master-weight gradients with ONE launch.  Groups follow the forward order in chunks of ~`group_bytes`, and a group's autograd
node is created where its first layer runs, so in backward it fires as soon as the last of its layers has produced its weight
gradient -- DDP's bucket all-reduces keep overlapping the rest of backward (one group for the whole network would hold every
gradient back until the end).

Only the `Trainer` turns a scaler on, and only for the duration of one step (`with scaler.step():`): the cache holds products of
the CURRENT parameter values, and the optimiser changes those at the end of the step.  Outside that window, or for a tensor
dtype the entry was not prepared for, layers take their own `weight * scale` path.
"""
import contextlib

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib

_CHUNK = 32768


class _Group:
    def __init__(self):
        self.entries = []          # _Entry
        self.outputs = None        # tuple of scaled weights of this step (aligned with `live`)
        self.live = None           # entries materialised this step
        self.tables = {}           # ("fwd" | "bwd", live signature) -> [host table, device table, block maps, key]


class _Entry:
    def __init__(self, module, scale, group):
        self.module, self.scale, self.group = module, float(scale), group
        self.dtype = None          # dtype of the tensor the layer multiplies with, learnt from its first call
        self.gain = 1.0


class _ScaleGroup(Function):
    @staticmethod
    def forward(ctx, scaler, group, *params):
        live = group.live
        outs = [torch.empty_like(p, dtype=e.dtype, memory_format=torch.preserve_format) for p, e in zip(params, live)]
        scaler._launch(group, "fwd", [(p, o, e.scale * e.gain) for p, o, e in zip(params, outs, live)])
        ctx.scaler, ctx.group, ctx.live = scaler, group, live
        ctx.shapes = [(p.shape, p.stride()) for p in params]
        ctx.save_for_backward(*params)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        params = ctx.saved_tensors
        rows, outs = [], []
        for p, g, e in zip(params, grads, ctx.live):
            if g is None:
                outs.append(None)
                continue
            if g.stride() != p.stride() or g.dtype not in (torch.float32, torch.bfloat16):   # flat kernel: same memory order as p
                g = torch.empty_like(p, dtype=g.dtype if g.dtype == torch.bfloat16 else torch.float32).copy_(g)
            o = torch.empty_like(p, memory_format=torch.preserve_format)
            rows.append((g, o, e.scale * e.gain))
            outs.append(o)
        if rows:
            ctx.scaler._launch(ctx.group, "bwd", rows)
        return (None, None) + tuple(outs)


class WeightScaler:
    def __init__(self, layers, group_bytes=24 << 20):
        """layers: [(module, scale)] in FORWARD order; module.weight is the fp32 master parameter."""
        self.groups, self.by_module = [], {}
        self.active = False
        for module, _ in layers:
            w = module.weight
            _lib.require_cuda(w)
            if w.dtype != torch.float32 or not (w.is_contiguous() or w.is_contiguous(memory_format=torch.channels_last)):
                raise RuntimeError("WeightScaler: master weights must be dense fp32 tensors")
        for members in plan_groups([m.weight.numel() * 4 for m, _ in layers], group_bytes):
            group = _Group()
            for i in members:
                module, scale = layers[i]
                e = _Entry(module, scale, group)
                group.entries.append(e)
                self.by_module[module] = e
            self.groups.append(group)
        self.device = layers[0][0].weight.device if layers else None

    @contextlib.contextmanager
    def step(self):
        """Scope of ONE training step: products are cached per group inside it and dropped at its end."""
        self.active = True
        try:
            yield self
        finally:
            self.active = False
            for g in self.groups:
                g.outputs = g.live = None

    def get(self, module, dtype, gain=1.0):
        """-> module.weight * scale * gain as `dtype`, or None (the layer then computes it itself)."""
        if not self.active or not torch.is_grad_enabled():
            return None
        e = self.by_module.get(module)
        if e is None:
            return None
        if e.dtype is None:                    # first sighting: learn what this layer multiplies with
            if dtype in (torch.float32, torch.bfloat16):
                e.dtype, e.gain = dtype, float(gain)
            return None
        if e.dtype != dtype or e.gain != float(gain):
            return None
        g = e.group
        if g.outputs is None:
            g.live = [x for x in g.entries if x.dtype is not None]
            g.outputs = _ScaleGroup.apply(self, g, *[x.module.weight for x in g.live])
        for x, o in zip(g.live, g.outputs):
            if x is e:
                return o
        return None

    # ---- one multi-tensor launch ----------------------------------------------------------------------------------------
    def _launch(self, group, kind, rows):
        """rows: [(src tensor, dst tensor, scale)].  The pointer table is rebuilt whenever a pointer moved (fresh tensors every
        eager step; fixed addresses inside a captured graph's pool) and shipped with one small pinned-memory copy."""
        import struct
        sig = (kind, tuple(int(s.numel()) for s, _, _ in rows))
        slot = group.tables.get(sig)
        if slot is None:
            bt, bc = [], []
            for ti, (s, _, _) in enumerate(rows):
                for c in range((s.numel() + _CHUNK - 1) // _CHUNK):
                    bt.append(ti)
                    bc.append(c)
            dev = rows[0][0].device
            slot = {"host": torch.zeros(len(rows) * 32, dtype=torch.uint8).pin_memory(),
                    "dev": torch.zeros(len(rows) * 32, dtype=torch.uint8, device=dev),
                    "bt": torch.tensor(bt, dtype=torch.int32, device=dev), "bc": torch.tensor(bc, dtype=torch.int32, device=dev),
                    "blocks": len(bt), "key": None}
            group.tables[sig] = slot
        key = tuple((s.data_ptr(), d.data_ptr()) for s, d, _ in rows)
        if key != slot["key"]:
            blob = b"".join(struct.pack("<QQqfi", s.data_ptr(), d.data_ptr(), s.numel(), float(sc),
                                        _lib.dtype_code(s) | (_lib.dtype_code(d) << 8)) for s, d, sc in rows)
            _lib.ship_table(slot, torch.frombuffer(bytearray(blob), dtype=torch.uint8))
            slot["key"] = key
        with torch.cuda.device(rows[0][0].device):
            _lib.check(_lib.load().gg_scale_cast_multi(slot["dev"].data_ptr(), slot["bt"].data_ptr(), slot["bc"].data_ptr(),
                                                       slot["blocks"], _CHUNK, _lib.stream()), "gg_scale_cast_multi")


def plan_groups(sizes, group_bytes):
    """Consecutive layers (forward order) -> groups of at most ~`group_bytes` of parameters; a layer larger than that is a
    group of its own.  -> list of index lists covering range(len(sizes)) in order."""
    groups, cur, acc = [], [], 0
    for i, sz in enumerate(sizes):
        if cur and acc + sz > group_bytes:
            groups.append(cur)
            cur, acc = [], 0
        cur.append(i)
        acc += sz
    if cur:
        groups.append(cur)
    return groups


def equalized_layers(network):
    """[(module, scale)] of every trainable EqualConv2d / EqualLinear of `network` in registration (= forward) order."""
    from ..stylegan2.networks import EqualConv2d, EqualLinear
    return [(m, float(m.scale)) for m in network.modules()
            if isinstance(m, (EqualConv2d, EqualLinear)) and m.weight.requires_grad]
