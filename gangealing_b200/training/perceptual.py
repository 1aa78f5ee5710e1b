"""Perceptual distance on VGG-16 features -- host-side mirror of the reference's `LPIPS` (models/losses/lpips.py:125-231)
in the two configurations its training script builds (lpips.py:13-22):

  'vgg_ssl' : LPIPS(net='vgg', lpips=False, pnet_rand=True, pretrained_weights=simclr_vgg_phase150.pt) / 18
  'lpips'   : LPIPS(net='vgg')  -- the linearly calibrated metric (`lin0..lin4`, 1x1 convs without bias)

The module tree and state-dict keys are the reference's (`scaling_layer.{shift,scale}`, `net.slice{k}.{torchvision index}.*`,
`lin{k}.model.1.weight`), so `load_state_dict` of a reference LPIPS checkpoint works, and `pretrained_weights=` loads a
`torchvision.models.vgg16().features` state dict strictly, exactly like lpips_backbones.py:103-105.  The VGG convolutions
stay cuDNN; the front end (normalise / difference / weights / spatial mean) is this package's fused
kernel (csrc/lpips.cu).  Offline benchmarks use seeded random VGG weights.
"""
import torch
import torch.nn as nn


_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
_SLICE_ENDS = (4, 9, 16, 23, 30)  # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 in torchvision's layer numbering
_CHNS = (64, 128, 256, 512, 512)


def _vgg16_features():
    layers, c_in = [], 3
    for v in _VGG16_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(c_in, v, kernel_size=3, padding=1), nn.ReLU(inplace=False)]
            c_in = v
    return layers


class VGG16Slices(nn.Module):
    """`slice1..slice5`, each an nn.Sequential whose children carry torchvision's `features` indices
    (reference lpips_backbones.py:106-121)."""

    def __init__(self):
        super().__init__()
        feats = _vgg16_features()
        starts = (0,) + _SLICE_ENDS[:-1]
        for k, (a, b) in enumerate(zip(starts, _SLICE_ENDS), 1):
            seq = nn.Sequential()
            for idx in range(a, b):
                seq.add_module(str(idx), feats[idx])
            setattr(self, "slice%d" % k, seq)
        self.N_slices = 5
        for p in self.parameters():
            p.requires_grad = False

    def slices(self):
        return [getattr(self, "slice%d" % k) for k in range(1, self.N_slices + 1)]

    def load_features_state_dict(self, state_dict, strict=True):
        """Load a `torchvision.models.vgg16().features` state dict ('0.weight', '2.bias', ...), strictly by default."""
        own = {}
        for k, seq in enumerate(self.slices(), 1):
            for idx, layer in seq.named_children():
                for name, _ in layer.named_parameters():
                    own["%s.%s" % (idx, name)] = "slice%d.%s.%s" % (k, idx, name)
        unexpected = [k for k in state_dict if k not in own]
        if strict and unexpected:
            raise RuntimeError("unexpected VGG16 feature keys (the reference slices stop at relu5_3): %s" % unexpected[:4])
        return self.load_state_dict({own[k]: v for k, v in state_dict.items() if k in own}, strict=strict)

    def forward(self, x, ops=None):
        """ops: the sm_100a op set -> every Conv2d + ReLU pair runs as a bias-free cuDNN convolution followed by ONE
        hand-written bias+ReLU pass (`fused_leaky_relu(x, bias, negative_slope=0, scale=1)` = relu(x + b), the channels-last
        streaming kernel of csrc/nhwc.cu) instead of cuDNN's separate broadcast bias-add kernel + ATen's clamp (two passes
        over the feature map; 1.2 ms of a 25 ms bf16 step, profiles/r02_step_b32_bf16_launches_v1.txt); at a slice boundary
        (ReLU -> tap -> MaxPool2d of the next slice) the bias+ReLU pass also emits the pooled map and the backward of
        pool + gradient add + ReLU is one pass (op/vgg_pool.py).  None: plain modules."""
        outs = []
        slices = self.slices()
        pooled = None          # the NEXT slice's MaxPool2d output, when the fused boundary kernel already produced it
        for k, s in enumerate(slices):
            if ops is None or not x.is_cuda:
                x = s(x)
            else:
                layers = list(s.children())
                nxt = list(slices[k + 1].children()) if k + 1 < len(slices) else []
                pool_next = bool(nxt) and isinstance(nxt[0], nn.MaxPool2d) and _is_2x2_pool(nxt[0]) and hasattr(ops, "bias_relu_pool")
                i = 0
                if pooled is not None:            # this slice's leading MaxPool2d ran inside the previous slice's last kernel
                    x, pooled, i = pooled, None, 1
                while i < len(layers):
                    m = layers[i]
                    if isinstance(m, nn.Conv2d) and i + 1 < len(layers) and isinstance(layers[i + 1], nn.ReLU):
                        x = nn.functional.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)
                        bias = m.bias.float() if m.bias.dtype != torch.float32 else m.bias
                        if i + 2 == len(layers) and pool_next and ops.bias_relu_pool_supported(x):
                            # slice boundary: relu(x + b) (the tapped feature map) AND its 2x2 max-pool in one pass
                            x, pooled = ops.bias_relu_pool(x, bias)
                        else:
                            x = ops.fused_leaky_relu(x, bias, 0.0, 1.0)
                        i += 2
                    else:
                        x = m(x)
                        i += 1
            outs.append(x)
        return outs


def _is_2x2_pool(m):
    def pair(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    return (pair(m.kernel_size) == (2, 2) and pair(m.stride) == (2, 2) and pair(m.padding) == (0, 0) and pair(m.dilation) == (1, 1)
            and not m.ceil_mode and not m.return_indices)


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """A 1x1 convolution without bias (reference lpips.py:236-245); with dropout the conv is `model.1`."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class PerceptualLoss(nn.Module):
    """d(x, y) = sum_layers mean_hw sum_c w_c (f/|f| - g/|g|)^2 on ImageNet-style rescaled inputs, / divisor.
    `lpips=False`: w = 1 (the 'vgg_ssl' baseline, divisor 18); `lpips=True`: w = the `lin{k}` weights (divisor 1)."""

    def __init__(self, divisor=18.0, lpips=False, use_dropout=True, pretrained_weights=None, ops=None):
        super().__init__()
        self.ops = ops        # None = the sm_100a op set (resolved lazily); tests / CPU legs inject the oracle's
        self.scaling_layer = ScalingLayer()
        self.net = VGG16Slices()
        self.lpips = lpips
        self.L = 5
        if lpips:
            for k, c in enumerate(_CHNS):
                setattr(self, "lin%d" % k, NetLinLayer(c, use_dropout=use_dropout))
            self.lins = nn.ModuleList([getattr(self, "lin%d" % k) for k in range(self.L)])
        if pretrained_weights is not None:
            sd = pretrained_weights if isinstance(pretrained_weights, dict) else \
                torch.load(pretrained_weights, map_location="cpu")
            self.net.load_features_state_dict(sd, strict=True)
        self.divisor = divisor
        self.eval()
        for p in self.parameters():
            p.requires_grad = False

    # the registered-twice `lins` ModuleList mirrors the reference (lpips.py:162-168); keep its keys out of the state
    # dict comparison the same way the reference's own checkpoints do (they hold both `lin0.*` and `lins.0.*`)

    def forward(self, in0, in1):
        # channels_last: cuDNN's tensor-core kernels are NHWC-native; NCHW inputs cost a layout conversion around
        # every convolution (22% of the step in profiles/r01_step_launches_b8_summary.txt)
        cl = torch.channels_last if in0.is_cuda else torch.contiguous_format
        dt = self.net.slice1[0].weight.dtype   # bf16 when the Trainer runs BASELINE config 3
        ops = self.ops
        if ops is None:
            from ..opset import cuda_ops
            ops = cuda_ops()
        native = ops if getattr(ops, "name", None) == "sm_100a" else None
        if native is not None and in0.is_cuda and in0.shape == in1.shape and hasattr(ops, "feature_distance_stacked"):
            # ONE backbone pass over both images stacked along the batch (the reference runs the VGG twice, lpips.py:188):
            # half the launches, larger convolutions; the distance kernel reads the two halves of each stacked map and
            # writes both gradients into one stacked tensor
            feats = self.net(self.scaling_layer(torch.cat([in0, in1], 0)).to(dt).contiguous(memory_format=cl), native)
            val = 0
            for k, f in enumerate(feats):
                w = self.lins[k].model[-1].weight.reshape(-1) if self.lpips else None
                val = val + ops.feature_distance_stacked(f, w)
            return val / self.divisor
        f0 = self.net(self.scaling_layer(in0).to(dt).contiguous(memory_format=cl), native)
        f1 = self.net(self.scaling_layer(in1).to(dt).contiguous(memory_format=cl), native)
        val = 0
        for k, (a, b) in enumerate(zip(f0, f1)):
            # normalise, difference, (weights,) channel sum and spatial mean in one pass over both maps (csrc/lpips.cu)
            # on channels-last CUDA features; lpips.py:193-205, :226
            w = self.lins[k].model[-1].weight.reshape(-1) if self.lpips else None
            val = val + ops.feature_distance(a, b, w)
        return val / self.divisor


def get_perceptual_loss(device, seed=0, ops=None, kind="vgg_ssl", pretrained_weights=None, lpips_weights=None):
    """kind='vgg_ssl' (default; reference lpips.py:14-17) or 'lpips' (lpips.py:18-20).  Without `pretrained_weights`
    the VGG is seeded-random (offline benchmark); `lpips_weights`: a reference LPIPS state dict / path with `lin*` keys."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    if kind == "vgg_ssl":
        loss = PerceptualLoss(divisor=18.0, lpips=False, pretrained_weights=pretrained_weights, ops=ops)
    elif kind == "lpips":
        loss = PerceptualLoss(divisor=1.0, lpips=True, pretrained_weights=pretrained_weights, ops=ops)
        if lpips_weights is not None:
            sd = lpips_weights if isinstance(lpips_weights, dict) else torch.load(lpips_weights, map_location="cpu")
            loss.load_state_dict(sd, strict=False)
    else:
        raise NotImplementedError(kind)
    torch.random.set_rng_state(g)
    loss = loss.to(device)
    if torch.device(device).type == "cuda":
        loss = loss.to(memory_format=torch.channels_last)
    return loss
