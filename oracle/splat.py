"""CPU restatement of Gaussian forward splatting (test infrastructure -- see oracle/__init__.py).

The reference has NO CPU implementation of this op (functional.py:54-55 raises on CPU) and its host wrapper no
longer compiles (THC).  This file follows the CUDA kernel utils/splat2d_cuda/src/splat_gpu_impl.cu:60-94 and the
host post-processing splat_gpu.c:20-41 line by line.  Pinning: on the GPU box tests/test_splat_gpu.py compares
both this restatement and the product kernel against the reference kernel itself (oracle/_ref/libsplat_ref.so,
built from the reference's .cu by oracle/build_ref.py).  Without that library the parity of this op is unpinned.
"""
import numpy as np
import torch


def splat2d_ref(input, coordinates, values, sigma, soft_normalize=False, return_alpha=False):
    inp = input.detach().cpu().double().numpy()
    coords = coordinates.detach().cpu().numpy().astype(np.float32)
    vals = values.detach().cpu().numpy().astype(np.float32)
    sig = sigma.detach().cpu().numpy().astype(np.float32)
    n, c, h, w = inp.shape
    acc = np.zeros((n, c, h, w), dtype=np.float64)
    alpha = np.zeros((n, h, w), dtype=np.float64)
    touched = np.zeros((n, h, w), dtype=bool)
    for i in range(n):
        sd = np.float32(sig[i])
        length = np.float32(2) * sd                                  # :68
        norm = -np.float32(1) / (np.float32(2) * sd * sd)            # :71
        x, y = coords[i, :, 0], coords[i, :, 1]
        inb = (x >= 0) & (x < w) & (y >= 0) & (y < h)                # :76
        x, y, v = x[inb], y[inb], vals[i][inb]
        t = np.maximum(0, np.floor(y - length)).astype(np.int64)     # :78-81
        b = np.minimum(h - 1, np.ceil(y + length)).astype(np.int64)
        l = np.maximum(0, np.floor(x - length)).astype(np.int64)
        r = np.minimum(w - 1, np.ceil(x + length)).astype(np.int64)
        span = int(np.ceil(2 * length)) + 2
        for dy in range(span + 1):
            for dx in range(span + 1):
                py, px = t + dy, l + dx
                ok = (py <= b) & (px <= r)
                if not ok.any():
                    continue
                pyo, pxo = py[ok], px[ok]
                d2 = (pxo.astype(np.float32) - x[ok]) ** 2 + (pyo.astype(np.float32) - y[ok]) ** 2
                a = np.exp((norm * d2).astype(np.float32)).astype(np.float64)   # :36-39
                np.add.at(alpha[i], (pyo, pxo), a)                              # :87
                touched[i][pyo, pxo] = True
                for ch in range(c):
                    np.add.at(acc[i, ch], (pyo, pxo), a * v[ok][:, ch].astype(np.float64))  # :89-91
    den = alpha[:, None]
    if soft_normalize:
        den = np.maximum(den, 1.0)                                   # splat_gpu.c:37-39
    out = torch.from_numpy(((inp + acc) / (den + 1e-8)).astype(np.float32))    # :21, :40
    if return_alpha:
        return out, torch.from_numpy(alpha.astype(np.float32)), torch.from_numpy(touched)
    return out


def splat_points_ref(images, points, sigma, opacity, colors, alpha_channel=None, blend_alg="alpha", splat_fn=splat2d_ref):
    """The splat2d call-site contract, reference utils/vis_tools/helpers.py:134-194 (alpha blending branch):
    two splats (colours, soft-normalised alpha) onto zero canvases, then alpha compositing."""
    n, c, h, w = images.shape
    if alpha_channel is None:
        alpha_channel = torch.ones(points.shape[0], points.shape[1], 1, device=points.device)
    sig = torch.full((n,), float(sigma), device=points.device) if not torch.is_tensor(sigma) else sigma
    blank_img = torch.zeros(n, colors.shape[-1], h, w, device=images.device)
    blank_mask = torch.zeros(n, 1, h, w, device=images.device)
    prop_obj = splat_fn(blank_img, points, colors, sig, False)
    prop_mask = splat_fn(blank_mask, points, alpha_channel, sig, True) * opacity
    return prop_mask * prop_obj + (1 - prop_mask) * images
