"""CPU restatement of the warp-parameter heads' arithmetic (test infrastructure -- see oracle/__init__.py).

Follows reference models/spatial_transformers/warping_heads.py:
  SimilarityHead.make_affine_matrix :36-50, make_3x3 :52-56, matrix composition :120-123
  FlowHead.upsample_flow :180-193 (RAFT convex upsampling), flow composition :239-251, apply_affine :268-277
"""
import math

import torch
import torch.nn.functional as F


def similarity_matrix_ref(params):
    """params (N, 4*K) split as [rot | scale | shift_x | shift_y] (torch.split(params, K, dim=1), :118-119)
    -> (N, K, 2, 3):  rot = tanh(.)*pi, scale = exp(.)  (:41-48)."""
    k = params.shape[1] // 4
    rot, scale, sx, sy = torch.split(params, k, dim=1)
    rot = torch.tanh(rot) * math.pi
    scale = torch.exp(scale)
    c, s = torch.cos(rot), torch.sin(rot)
    m = torch.stack([scale * c, -scale * s, sx, scale * s, scale * c, sy], dim=2)
    return m.reshape(params.shape[0], k, 2, 3)


def compose_similarity_ref(base_warp, matrix):
    """base_warp @ [[matrix], [0, 0, 1]]  (:120-123, :52-56)."""
    if base_warp.dim() == 3:
        base_warp = base_warp.unsqueeze(1)
    last = torch.tensor([0.0, 0.0, 1.0]).reshape(1, 1, 1, 3).expand(matrix.shape[0], matrix.shape[1], 1, 3)
    return base_warp @ torch.cat([matrix, last], dim=2)


def upsample_flow_ref(flow, mask, s=8):
    """flow (N, H, W, 2), mask (N, 9*s*s, H, W) -> (N, s*H, s*W, 2): convex combination of the 3x3
    neighbourhood of s*flow with softmax(mask) weights (:180-193)."""
    n, h, w, _ = flow.shape
    weights = torch.softmax(mask.reshape(n, 9, s, s, h, w), dim=1)                    # (N, 9, sy, sx, H, W)
    padded = F.pad(s * flow.permute(0, 3, 1, 2), (1, 1, 1, 1))                         # (N, 2, H+2, W+2), zeros
    out = flow.new_zeros(n, 2, s, s, h, w)
    for k in range(9):                                                                 # unfold order: ky*3 + kx
        ky, kx = divmod(k, 3)
        nb = padded[:, :, ky:ky + h, kx:kx + w]                                        # (N, 2, H, W)
        out = out + weights[:, k].unsqueeze(1) * nb[:, :, None, None]
    # (N, 2, sy, sx, H, W) -> (N, H, sy, W, sx, 2) -> (N, sH, sW, 2)
    return out.permute(0, 4, 2, 5, 3, 1).reshape(n, s * h, s * w, 2)


def apply_affine_ref(matrix, grid):
    """[gx, gy, 1] @ matrix^T for every grid point (:268-277)."""
    gx, gy = grid[..., 0], grid[..., 1]
    m = matrix.reshape(-1, 1, 1, 2, 3)
    return torch.stack([m[..., 0, 0] * gx + m[..., 0, 1] * gy + m[..., 0, 2],
                        m[..., 1, 0] * gx + m[..., 1, 1] * gy + m[..., 1, 2]], dim=-1)


def identity_flow_ref(size_h, size_w):
    """FlowHead.initialize_flow (:172-178): F.affine_grid(identity) of the full-resolution flow."""
    return F.affine_grid(torch.eye(2, 3).unsqueeze(0), (1, 1, size_h, size_w), align_corners=False)


def flow_compose_ref(low_flow, mask, identity_flow, base_warp=None, alpha=None, s=8):
    """FlowHead.forward :239-244.  Returns (delta_flow, flow)."""
    delta = upsample_flow_ref(low_flow, mask, s)
    flow = identity_flow + delta
    if base_warp is not None:
        flow = apply_affine_ref(base_warp, flow)
    if alpha is not None:
        flow = identity_flow.lerp(flow, alpha[:, None, None, None])
    return delta, flow


def resize_flow_ref(flow, output_resolution):
    """:245-251: bilinear resize of the sampling grid itself (align_corners=False)."""
    return F.interpolate(flow.permute(0, 3, 1, 2), scale_factor=output_resolution / flow.size(2),
                         mode="bilinear").permute(0, 2, 3, 1)
