"""GPU: the cluster classifier (BASELINE config 5, second half) on the sm_100a op set against the reference-generated fixture
(tests/golden/classifier.npz) -- logits in the NCHW, channels-last fp32 and channels-last bf16 trunks, exact index outputs of
the inference helpers -- and one ClassifierTrainer iteration per storage type.  The classifier launches no kernel of its own:
its trunk is the similarity STN's.  First run on a B200: profiles/r02_gputest_classifier.txt (5 passed)."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle.make_golden import classifier_setup

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mods():
    from gangealing_b200.cluster_classifier import ResnetClassifier
    from gangealing_b200.stn import BilinearDownsample, get_stn
    from gangealing_b200.stylegan2 import Generator
    from gangealing_b200.training import DirectionInterpolator
    return dict(Generator=Generator, get_stn=get_stn, DirectionInterpolator=DirectionInterpolator,
                ResnetClassifier=ResnetClassifier, BilinearDownsample=BilinearDownsample)


@pytest.mark.parametrize("layout", ["nchw", "nhwc", "nhwc_bf16"])
def test_classifier_on_gpu_matches_reference_fixture(layout):
    blob = load_golden("classifier")
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        _, _, _, cls, _, clusters = classifier_setup(_mods())
        cls.to(DEV)
        if layout != "nchw":
            cls.to(memory_format=torch.channels_last)
            cls.channels_last = True
            cls.act_dtype = torch.bfloat16 if layout == "nhwc_bf16" else torch.float32
        x = blob["cls.x"].to(DEV)
        with torch.no_grad():
            logits = cls(x)
            assert logits.dtype == torch.float32
            assert_close(logits, blob["cls.logits"], rtol=5e-2 if layout == "nhwc_bf16" else 1e-3, what="logits " + layout)
            if layout != "nhwc_bf16":   # margins of the fixture's logits are >= 0.1: index outputs are exact in fp32
                assert torch.equal(cls.assign(x).cpu(), blob["cls.assign"])
                flipped, _, classes, flip = cls.run_flip(x)
                assert torch.equal(classes.cpu(), blob["cls.run_flip.classes"]) and torch.equal(flip.cpu(), blob["cls.run_flip.flip"])
                tiled, policy = cls.run_flip_cartesian(x)
                assert torch.equal(policy.cpu(), blob["cls.cartesian.policy"])
                assert torch.equal(tiled[..., ::8, 3::8].cpu(), blob["cls.cartesian.out"])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_classifier_trainer_step_on_gpu(dtype):
    """One train_cluster_classifier.py iteration on the sm_100a op set (config-5 shapes shrunk): finite loss, histograms sum to
    one, every classifier parameter moves through the fused Adam kernel, the clustering STN stays frozen."""
    from gangealing_b200.training import TrainConfig, Trainer
    from gangealing_b200.training.classifier_step import ClassifierTrainer
    cfg = TrainConfig(gen_size=128, flow_size=64, dim_latent=64, n_mlp=2, batch=4, inject=3, num_heads=2, flips=True, ndirs=2,
                      sample_from_full_res=True, padding_mode="reflection", dtype=dtype)
    tr = Trainer(cfg, DEV)
    ct = ClassifierTrainer(tr, cls_lr=1e-3)
    before = [p.detach().clone() for p in ct.module.parameters()]
    stn_before = [p.detach().clone() for p in tr.t_ema.parameters()]
    for _ in range(2):
        out = ct.step()
    assert torch.isfinite(out["cross_entropy"]).item()
    assert abs(sum(float(out["head_%d" % c]) for c in range(4)) - 1.0) < 1e-6
    assert abs(sum(float(out["pred_head_%d" % c]) for c in range(4)) - 1.0) < 1e-6
    assert all(not torch.equal(a, b) for a, b in zip(before, ct.module.parameters()))
    assert all(torch.equal(a, b) for a, b in zip(stn_before, tr.t_ema.parameters()))
