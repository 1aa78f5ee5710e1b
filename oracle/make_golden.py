"""Generate tests/golden/*.npz from the reference itself and pin the oracle against it.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
For every case the script (1) runs the REFERENCE's own CPU implementation on seeded inputs, (2) asserts
that the oracle restatement reproduces it (tight tolerance / exact where integer), (3) stores inputs and
reference outputs as a small fixture.  Fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import refimport, stylegan2_ops as so  # noqa: E402


def _save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote %-40s %7.1f KB" % (os.path.relpath(path, ROOT), os.path.getsize(path) / 1024))


def _close(a, b, tol, what):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err <= tol * max(1.0, ref), "%s: oracle deviates from the reference by %g" % (what, err)


# ---------------------------------------------------------------------------------------------------
UPFIRDN_CASES = [
    # name, shape, kernel spec, up, down, pad
    ("g_blur_9", (2, 3, 9, 9), "1331x4", 1, 1, (1, 1)),         # Generator blur after up-conv (networks.py:199-205)
    ("g_blur_65", (1, 2, 65, 65), "1331x4", 1, 1, (1, 1)),
    ("stn_blur_pad22", (1, 2, 32, 32), "1331", 1, 1, (2, 2)),   # ResBlock conv2 blur (networks.py:606-611)
    ("stn_blur_pad11", (1, 2, 32, 32), "1331", 1, 1, (1, 1)),   # ResBlock skip blur
    ("stn_blur_128", (1, 1, 128, 128), "1331", 1, 1, (1, 1)),   # 127-wide output: ragged strips
    ("blur_bwd_pad22", (1, 2, 64, 64), "1331x4", 1, 1, (2, 2)),
    ("rgb_up2", (2, 3, 8, 8), "1331x4", 2, 1, (2, 1)),          # Upsample (networks.py:28-46)
    ("rgb_up2_bwd_dn2", (2, 3, 16, 16), "1331x4", 1, 2, (1, 1)),
    ("down2", (1, 2, 16, 16), "1331", 1, 2, (1, 1)),            # Downsample (networks.py:49-67)
    ("k5_up2_dn3", (1, 2, 11, 13), "rand5", 2, 3, (3, 2)),      # generic path
    ("k3_asym", (1, 2, 12, 10), "rand3", 1, 1, (1, 1)),         # flip matters
    ("k2", (1, 1, 7, 9), "rand2", 1, 1, (0, 1)),
    ("neg_pad", (1, 2, 20, 20), "1331", 1, 1, (-1, 2)),         # negative pad crops
    ("k43_rect", (1, 1, 15, 15), "rand43", 1, 1, (2, 1)),
    ("tall_bands", (1, 1, 300, 40), "1331", 1, 1, (2, 2)),      # several bands per plane
]


def _kernel(spec, gen):
    if spec.startswith("1331"):
        k = so.make_kernel([1, 3, 3, 1])
        return k * 4 if spec.endswith("x4") else k
    if spec == "rand43":
        return torch.randn(4, 3, generator=gen)
    n = int(spec[4:])
    return torch.randn(n, n, generator=gen)


def gen_upfirdn2d(ref_models):
    from models.stylegan2.op.upfirdn2d import upfirdn2d as ref_upfirdn2d
    out = {}
    for i, (name, shape, spec, up, down, pad) in enumerate(UPFIRDN_CASES):
        gen = torch.Generator().manual_seed(1000 + i)
        x = torch.randn(*shape, generator=gen)
        k = _kernel(spec, gen)
        y_ref = ref_upfirdn2d(x, k, up=up, down=down, pad=pad)
        y_or = so.upfirdn2d_ref(x, k, up=up, down=down, pad=pad)
        _close(y_or, y_ref, 1e-6, "upfirdn2d/" + name)
        out[name + ".x"] = x
        out[name + ".k"] = k
        out[name + ".cfg"] = np.array([up, down, pad[0], pad[1]])
        out[name + ".y"] = y_ref
    _save("upfirdn2d", **out)


def gen_fused_act(ref_models):
    from models.stylegan2.op.fused_act import fused_leaky_relu as ref_flr
    out = {}
    shapes = [(3, 5, 4, 4), (4, 7), (2, 6, 3, 5), (1, 4, 33, 31), (2, 8, 16, 16)]
    for i, shape in enumerate(shapes):
        gen = torch.Generator().manual_seed(2000 + i)
        x = torch.randn(*shape, generator=gen, requires_grad=True)
        b = torch.randn(shape[1], generator=gen, requires_grad=True)
        y_ref = ref_flr(x, b)  # CPU branch: leaky_relu(x + b, 0.2) * sqrt(2)
        g = torch.randn(*shape, generator=gen)
        gx_ref, gb_ref = torch.autograd.grad(y_ref, [x, b], g)
        y_or = so.fused_leaky_relu_ref(x.detach(), b.detach())
        gx_or, gb_or = so.fused_leaky_relu_backward_ref(g, y_or)
        _close(y_or, y_ref.detach(), 1e-6, "fused_leaky_relu fwd %s" % (shape,))
        _close(gx_or, gx_ref, 1e-6, "fused_leaky_relu gx %s" % (shape,))
        _close(gb_or, gb_ref, 1e-5, "fused_leaky_relu gb %s" % (shape,))
        tag = "case%d" % i
        out[tag + ".x"], out[tag + ".b"], out[tag + ".g"] = x.detach(), b.detach(), g
        out[tag + ".y"], out[tag + ".gx"], out[tag + ".gb"] = y_ref.detach(), gx_ref, gb_ref
    _save("fused_act", **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref_models = refimport.import_reference()
    gen_upfirdn2d(ref_models)
    gen_fused_act(ref_models)
    for extra in EXTRA_GENERATORS:
        extra(ref_models)


EXTRA_GENERATORS = []

if __name__ == "__main__":
    main()
