"""ResnetClassifier -- host-side mirror of reference models/cluster_classifier.py (the cluster / flip predictor of the
clustering models, BASELINE config 5: `train_cluster_classifier.py`, `ComposedSTN.load_average_warp_and_flip`).

Same module tree and `state_dict` keys as the reference (`input_downsample`, `convs.*`, `final_conv.*`, `to_logits.*`): the
trunk IS the similarity STN's trunk (reference train_cluster_classifier.py:189-193 initialises it from `t_ema.stns[0]`), so it
runs on the same sm_100a kernels as the STN -- channels-last Blur (`gg_blur_nhwc` mode 0), bias+lrelu forward / backward
(`gg_noise_bias_act_nhwc`, `gg_bias_act_backward_nhwc`), tent down-sampling (`gg_tent_downsample_*`) -- and needs no kernel
of its own.  `accuracy` is reference models/__init__.py:34-42.
"""
import math

import torch
import torch.nn as nn

from .opset import cuda_ops
from .stn.sampling import BilinearDownsample
from .stylegan2.networks import ConvLayer, EqualLinear, ResBlock, channel_table

_IGNORED_KEYS = ("input_downsample.kernel_horz", "input_downsample.kernel_vert")


class ResnetClassifier(nn.Module):
    """image (N, 3, S, S) -> logits (N, num_heads); num_heads = clusters x (1 + flips)   (reference :8-49)."""

    def __init__(self, size, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], num_heads=1, supersize=None, ops=None):
        super().__init__()
        self.ops = ops if ops is not None else cuda_ops()
        self.stn_in_size = size
        self.num_heads = num_heads
        if supersize is not None:
            self.input_downsample = BilinearDownsample(supersize // size, 3, ops=ops)
        self.channels_last = False        # trunk activations NHWC (CUDA only), as SpatialTransformer.channels_last
        self.act_dtype = torch.float32    # storage type of the trunk's activations; the logits are always fp32
        channels = channel_table(channel_multiplier)
        convs = [ConvLayer(3, int(channels[size]), 1, ops=ops)]
        in_channel = channels[size]
        for i in range(int(math.log(size, 2)), 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(int(in_channel), int(out_channel), blur_kernel, ops=ops))
            in_channel = out_channel
        self.convs = nn.Sequential(*convs)
        self.final_conv = ConvLayer(int(in_channel), channels[4], 3, ops=ops)
        self.to_logits = EqualLinear(channels[4] * 4 * 4, num_heads, activation="fused_lrelu", ops=ops)

    def forward(self, input):
        if input.size(-1) > self.stn_in_size:
            input = self.input_downsample(input)
        if self.channels_last and input.is_cuda:
            input = input.to(self.act_dtype).contiguous(memory_format=torch.channels_last)
        feat = self.final_conv(self.convs(input))
        if feat.dtype != torch.float32:
            feat = feat.float()
        return self.to_logits(feat.reshape(feat.shape[0], -1))   # logical (C, H, W) order in either layout

    # ---- the reference's inference helpers (:51-100): argmax over the logits, heads [0, K) unflipped, [K, 2K) mirrored ----
    def assign(self, input, ignore_flips=False):
        classes = self.forward(input).argmax(dim=1)
        return classes % (self.num_heads // 2) if ignore_flips else classes

    @staticmethod
    def _mirror_where(flip, images):
        return torch.where(flip.reshape(-1, *([1] * (images.dim() - 1))), images.flip(images.dim() - 1), images)

    def run(self, input, target_cluster, return_flip_indices=False):
        """Keep the images assigned to `target_cluster` (either orientation), mirrored where the flip head won."""
        half = self.num_heads // 2
        preds = self.forward(input)
        classes = preds.argmax(dim=1)
        (keep,) = torch.where((classes % half) == target_cluster)
        flip = (classes[keep] >= half).reshape(keep.size(0), 1, 1, 1)
        kept = self._mirror_where(flip, input[keep])
        return (kept, preds[keep], flip, keep) if return_flip_indices else (kept, preds[keep])

    def run_flip(self, input):
        preds = self.forward(input)
        classes = preds.argmax(dim=1)
        flip = classes >= self.num_heads // 2
        return self._mirror_where(flip, input), preds, classes, flip

    def run_flip_target(self, input, target_cluster):
        pair = self.forward(input)[:, [target_cluster, target_cluster + self.num_heads // 2]]
        flip = pair.argmax(dim=1) == 1
        return self._mirror_where(flip, input), flip

    def run_flip_cartesian(self, input):
        """Every image paired with every cluster, each pair in the orientation the classifier prefers for that cluster."""
        half, n = self.num_heads // 2, input.size(0)
        flip = self.forward(input).view(n, 2, half).argmax(dim=1) == 1                  # (N, K)
        tiled = input.unsqueeze(1).repeat(1, half, 1, 1, 1)
        tiled = torch.where(flip.reshape(n, half, 1, 1, 1), tiled.flip(4), tiled)
        policy = torch.eye(half, device=input.device).repeat(n, 1)
        return tiled.view(n * half, *input.shape[1:]), policy

    def load_state_dict(self, state_dict, strict=True):
        return super().load_state_dict({k: v for k, v in state_dict.items() if k not in _IGNORED_KEYS}, False)


@torch.no_grad()
def accuracy(predictions, gt_probabilities, k=1):
    """"Reverse" top-k accuracy: is the classifier's argmax among the k best classes of the ground truth?"""
    top_pred = predictions.argmax(dim=1, keepdim=True)
    top_gt = gt_probabilities.topk(k=k, dim=1).indices
    return (top_pred == top_gt).any(dim=1).float().mean()
