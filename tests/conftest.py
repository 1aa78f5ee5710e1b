import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """tests/golden/<name>.npz -> dict of torch tensors (fixtures written by oracle/make_golden.py)."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as data:
        return {k: torch.from_numpy(data[k]) for k in data.files}


def golden_cases(blob):
    names = []
    for k in blob:
        n = k.rsplit(".", 1)[0]
        if n not in names:
            names.append(n)
    return names


def assert_close(actual, expected, rtol=1e-3, atol=None, what=""):
    """The north-star tolerance: 1e-3 relative (to the tensor's magnitude) in fp32."""
    actual = actual.detach().float().cpu()
    expected = expected.detach().float().cpu()
    assert actual.shape == expected.shape, "%s: shape %s vs %s" % (what, tuple(actual.shape), tuple(expected.shape))
    scale = expected.abs().max().item()
    tol = rtol * max(scale, 1e-6) if atol is None else atol
    err = (actual - expected).abs().max().item() if actual.numel() else 0.0
    assert err <= tol, "%s: max abs err %.3e > %.3e (ref magnitude %.3e)" % (what, err, tol, scale)
