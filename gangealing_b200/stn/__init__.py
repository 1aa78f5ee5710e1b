"""Spatial Transformer side of the hot path (mirror of reference models/spatial_transformers/)."""
from .sampling import BilinearDownsample, MipmapWarp, Warp, grid_sample_bilinear
from .flow import apply_affine, flow_compose, upsample_flow

__all__ = ["BilinearDownsample", "MipmapWarp", "Warp", "grid_sample_bilinear", "apply_affine", "flow_compose",
           "upsample_flow"]
from .heads import FlowHead, SimilarityHead
from .transformer import ComposedSTN, SpatialTransformer, get_stn, total_variation_loss
__all__ += ["FlowHead", "SimilarityHead", "ComposedSTN", "SpatialTransformer", "get_stn", "total_variation_loss"]
