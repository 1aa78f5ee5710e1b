"""Perceptual distance on VGG-16 features -- the `LPIPS(net='vgg', lpips=False, pnet_rand=True)` baseline the
reference's 'vgg_ssl' loss builds (models/losses/lpips.py:13-17, :181-223; backbone lpips_backbones.py:98-140),
divided by 18 like the reference.  OUT OF SCOPE for hand-written kernels (plain cuDNN convs);
it exists so the training step is complete.  No weights are downloaded: random (seeded) initialisation stands in
for the SimCLR checkpoint, exactly as BASELINE.md section 3 prescribes for the offline benchmark."""
import torch
import torch.nn as nn


_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
_SLICE_ENDS = (4, 9, 16, 23, 30)  # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 in torchvision's layer numbering


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
    def __init__(self):
        super().__init__()
        feats = _vgg16_features()
        starts = (0,) + _SLICE_ENDS[:-1]
        self.slices = nn.ModuleList([nn.Sequential(*feats[a:b]) for a, b in zip(starts, _SLICE_ENDS)])
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        outs = []
        for s in self.slices:
            x = s(x)
            outs.append(x)
        return outs


class PerceptualLoss(nn.Module):
    """d(x, y) = sum_layers mean_hw sum_c (f/|f| - g/|g|)^2 on ImageNet-style rescaled inputs, / 18."""

    def __init__(self, divisor=18.0, ops=None):
        super().__init__()
        self.ops = ops        # None = the sm_100a op set (resolved lazily); tests / CPU legs inject the oracle's
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])
        self.net = VGG16Slices()
        self.divisor = divisor
        self.eval()

    @staticmethod
    def _unit(feat, eps=1e-10):
        return feat / (torch.sqrt(torch.sum(feat ** 2, dim=1, keepdim=True)) + eps)

    def forward(self, in0, in1):
        # channels_last: cuDNN's tensor-core kernels are NHWC-native; NCHW inputs cost a layout conversion around
        # every convolution (22% of the step in profiles/r01_step_launches_b8_summary.txt)
        cl = torch.channels_last if in0.is_cuda else torch.contiguous_format
        f0 = self.net(((in0 - self.shift) / self.scale).contiguous(memory_format=cl))
        f1 = self.net(((in1 - self.shift) / self.scale).contiguous(memory_format=cl))
        ops = self.ops
        if ops is None:
            from ..opset import cuda_ops
            ops = cuda_ops()
        val = 0
        for a, b in zip(f0, f1):
            # normalise, difference, channel sum and spatial mean in one pass over both maps (csrc/lpips.cu) on
            # channels-last CUDA features; the same formula with tensor ops elsewhere (lpips.py:193-205, :226)
            val = val + ops.feature_distance(a, b)
        return val / self.divisor


def get_perceptual_loss(device, seed=0, ops=None):
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    loss = PerceptualLoss(ops=ops)
    torch.random.set_rng_state(g)
    loss = loss.to(device)
    if torch.device(device).type == "cuda":
        loss = loss.to(memory_format=torch.channels_last)
    return loss
