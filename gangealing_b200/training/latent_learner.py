"""DirectionInterpolator -- learned truncation of the aligned target's latent (host-side mirror of reference
models/latent_learner.py:25-82; buffers `directions`, `lat_mean`, parameter `coefficients`).  The PCA fit and
k-means++ initialisers of the reference are init-time utilities outside the hot path;
offline benchmarks use seeded random `directions`/`lat_mean` buffers."""
import torch
import torch.nn as nn


class DirectionInterpolator(nn.Module):
    def __init__(self, pca_path, n_comps, inject_index, n_latent, num_heads=1, initializer=None, dim_latent=512):
        super().__init__()
        if pca_path is not None:
            import numpy as np
            with np.load(pca_path) as data:
                self.register_buffer("lat_mean", torch.from_numpy(data["lat_mean"]))
                self.register_buffer("directions", torch.from_numpy(data["lat_comp"].squeeze(axis=1))[:n_comps])
        else:
            self.register_buffer("directions", torch.randn(n_comps, dim_latent))
            self.register_buffer("lat_mean", torch.randn(1, dim_latent))
        if initializer is None:
            initializer = torch.zeros(num_heads, n_comps)
        self.coefficients = nn.Parameter(initializer.detach().clone())
        self.n_latent, self.inject_index, self.num_heads = n_latent, inject_index, num_heads

    def forward(self, styled_latent, psi=None, lat_mean=None, pca=None, unfold=False):
        if pca is not None:
            return self.assign_buffers(pca)
        return self.interpolate(styled_latent, psi, lat_mean, unfold)

    def interpolate(self, styled_latent, psi, lat_mean=None, unfold=False):
        assert len(styled_latent) == 1
        w = styled_latent[0]
        n = w.size(0)
        mean = self.lat_mean if lat_mean is None else lat_mean
        target = (mean + self.coefficients @ self.directions).repeat(n, 1)          # (N*K, D)
        w = w.repeat_interleave(self.num_heads, dim=0)
        truncated = target.lerp(w, psi).unsqueeze(1).repeat(1, self.inject_index, 1)
        fixed = w.unsqueeze(1).repeat(1, self.n_latent - self.inject_index, 1)
        out = torch.cat([truncated, fixed], dim=1)
        if unfold:
            out = out.reshape(n, self.num_heads, self.n_latent, out.size(-1))
        return [out]

    @torch.no_grad()
    def assign_buffers(self, pca):
        dev = self.directions.device
        self.register_buffer("directions", torch.from_numpy(pca.pca.components_).float().to(dev))
        self.register_buffer("lat_mean", torch.from_numpy(pca.pca.mean_[None]).float().to(dev))
