"""CPU: the sampler restatement (explicit index arithmetic) reproduces the reference-generated fixtures."""
import torch

from conftest import assert_close, golden_cases, load_golden
from oracle import sampling as S


def test_mipmap_warp_oracle_matches_reference_fixtures():
    blob = load_golden("mipmap_warp")
    names = golden_cases(blob)
    assert len(names) >= 8
    for name in names:
        mode = S.PAD_MODES[int(blob[name + ".mode"])]
        x = blob[name + ".x"].clone().requires_grad_(True)
        grid = blob[name + ".grid"].clone().requires_grad_(True)
        y, aux = S.mipmap_warp_ref(x, grid, 3.5, 0.0, mode, return_aux=True)
        assert_close(y, blob[name + ".y"], rtol=5e-6, what=name)
        gx, gg = torch.autograd.grad(y, [x, grid], blob[name + ".go"])
        assert_close(gx, blob[name + ".gx"], rtol=2e-5, what=name + " gx")
        assert_close(gg, blob[name + ".ggrid"], rtol=2e-4, what=name + " ggrid")
        assert_close(aux["levels"], blob[name + ".levels"], rtol=1e-6, what=name + " levels")
        assert_close(S.warp_ref(blob[name + ".x"], blob[name + ".grid"], mode), blob[name + ".warp_y"], rtol=5e-6)


def test_level_indices_are_integers_in_range():
    blob = load_golden("mipmap_warp")
    for name in golden_cases(blob):
        _, aux = S.mipmap_warp_ref(blob[name + ".x"], blob[name + ".grid"], 3.5, 0.0, "border", return_aux=True)
        assert aux["level_0"].min() >= 0 and aux["level_1"].max() <= 3
        assert aux["num_levels"] == int(aux["level_1"].max()) + 1


def test_bilinear_downsample_oracle():
    blob = load_golden("bilinear_downsample")
    for stride in (2, 4):
        assert_close(S.bilinear_downsample_ref(blob["x"], stride), blob["s%d.y" % stride], rtol=1e-6)


def test_affine_grid_restatement():
    import torch.nn.functional as F
    theta = torch.tensor([[[1.2, -0.3, 0.1], [0.3, 1.2, -0.2]]])
    assert_close(S.affine_grid_ref(theta, (1, 3, 5, 7)), F.affine_grid(theta, (1, 3, 5, 7), align_corners=False), rtol=1e-6)
