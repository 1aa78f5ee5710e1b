"""Import the read-only reference checkout (CPU code paths only) -- used by oracle/make_golden.py and by
the container-only pinning tests.  The reference JIT-builds its CUDA extensions at import time
(models/stylegan2/op/upfirdn2d.py:9-16, fused_act.py:10-17); that is stubbed out here because only the
reference's CPU branches are executed (upfirdn2d.py:146-149, fused_act.py:87-94)."""
import os
import sys

REFERENCE_ROOT = os.environ.get("GG_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


_done = False


def import_reference():
    """Returns the reference's `models` package (imported from REFERENCE_ROOT)."""
    global _done
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    if not _done:
        import torch.utils.cpp_extension as cpp_ext

        class _NoNative:  # any attribute access means a CUDA path was taken by mistake
            def __getattr__(self, name):
                raise RuntimeError("reference native extension is stubbed out (CPU paths only)")

        cpp_ext.load = lambda *a, **k: _NoNative()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _done = True
    import models  # noqa: F401  (the reference's package)
    return sys.modules["models"]
