"""CPU: libgg_b200.so loads without a GPU and exports exactly the symbols include/gg_b200.h declares."""
import os
import re

from conftest import ROOT
from gangealing_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "gg_b200.h")).read()
    return sorted(set(re.findall(r"GG_API\s+[\w\s\*]+?\b(gg_\w+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    declared = _declared()
    assert "gg_upfirdn2d" in declared and "gg_fused_bias_act" in declared
    dll = _lib.load()
    for name in declared:
        assert hasattr(dll, name), "libgg_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"


def test_version_and_error_string():
    dll = _lib.load()
    assert dll.gg_version() >= 1
    assert isinstance(dll.gg_last_error(), bytes)


def test_bad_arguments_are_reported_not_fatal():
    dll = _lib.load()
    # no device work is reached: argument validation happens first
    rc = dll.gg_fused_bias_act(None, None, None, None, 0, 3, 0, 0.2, 1.0, 16, 1, 0, None)
    assert rc == -1 and b"null" in dll.gg_last_error()
    rc = dll.gg_fused_bias_act(None, None, None, None, 0, 7, 0, 0.2, 1.0, 16, 1, 0, None)
    assert rc < 0
    rc = dll.gg_upfirdn2d(None, None, None, 3, 1, 4, 4, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, None)
    assert rc < 0  # filter larger than input / unsupported dtype
    assert dll.gg_bias_act_backward_workspace(2, 3, 4096) == 2 * 3 * 4


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from gangealing_b200.op import fused_leaky_relu, upfirdn2d
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        upfirdn2d(torch.zeros(1, 1, 8, 8), torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        fused_leaky_relu(torch.zeros(1, 2, 4, 4), torch.zeros(2))
