"""gangealing_b200 -- Blackwell (sm_100a) kernels for GANgealing's per-step hot path.

Scope: the frozen StyleGAN2 generator forward that synthesises each training
pair and the Spatial Transformer that warps it, behind the reference's op-level Python API
(`upfirdn2d`, `fused_leaky_relu`/`FusedLeakyReLU`, antialiased `grid_sample` (`MipmapWarp`), flow
composition (`apply_affine`, convex flow upsampling) and `splat2d`).  All device work goes through
the C ABI of libgg_b200.so (include/gg_b200.h); there is no CPU fallback -- calling an op without the
built library or with CPU tensors raises.
"""
__version__ = "0.1.0"
