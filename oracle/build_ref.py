"""Compile the REFERENCE's own CUDA kernels, from the sources where they lie under /root/reference, into
oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).  Test infrastructure: on the B200 the parity
tests and tools/opbench.py load these to compare against -- and time -- the reference's kernels recompiled
for sm_100a ("the kernel bar").  No reference source is copied into the repo.

  upfirdn2d_ref.so / fused_ref.so : models/stylegan2/op/{upfirdn2d,fused_bias_act}{.cpp,_kernel.cu}, built with
                                    torch.utils.cpp_extension (pybind modules, need the image's torch to load)
  libsplat_ref.so                 : utils/splat2d_cuda/src/splat_gpu_impl.cu alone (plain nvcc -shared); its
                                    extern "C" SplatForwardGpu (splat_gpu_impl.cuh:11-22) is called through ctypes.
                                    The reference's host wrapper splat_gpu.c includes <THC/THC.h>, which modern
                                    torch no longer ships, so it is not built; oracle/splat.py restates its
                                    clone / zeros / clamp / divide (splat_gpu.c:20-41).
  refpy/                          : the reference's own Python networks (models/**, utils/{__init__,distributed,
                                    download}.py) BYTE-COMPILED (py_compile, sourceless .pyc -- a compiled output
                                    like the .so files, no source text enters the repo) so that the UNMODIFIED
                                    reference Generator / get_stn / gangealing_loss can execute on the GPU box above
                                    gangealing_b200.compat (tests/test_reference_dropin_gpu.py: the "drops in
                                    unchanged" claim of SURVEY.md 8(b)).  models/stylegan2/op is NOT compiled (it is
                                    the kernel boundary and JIT-builds CUDA at import); antialiased_sampling.py is
                                    compiled only so that tools/opbench.py --stn can time the reference's MipmapWarp.
  refpy_cpu/                      : the same byte-compiled tree WITH models/stylegan2/op: the complete reference
                                    network code for the host-CPU arm of bench.py (oracle/reference_step.py stubs the
                                    import-time JIT build, the reference then takes its own native CPU branches).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("GG_REFERENCE_ROOT", "/root/reference")


def build(verbose=False):
    if not os.path.isdir(REF):
        raise RuntimeError("reference checkout not found at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    built = []
    # --- splat: plain nvcc
    lib = os.path.join(OUT, "libsplat_ref.so")
    src = os.path.join(REF, "utils", "splat2d_cuda", "src", "splat_gpu_impl.cu")
    if not os.path.exists(lib):
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
               "-I", os.path.dirname(src), src, "-o", lib]
        subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    built.append(lib)
    # --- StyleGAN2 ops: pybind extensions through torch
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    opdir = os.path.join(REF, "models", "stylegan2", "op")
    for name, files in (("upfirdn2d_ref", ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"]),
                        ("fused_ref", ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"])):
        target = os.path.join(OUT, name + ".so")
        if not os.path.exists(target):
            bdir = os.path.join(OUT, "build_" + name)
            os.makedirs(bdir, exist_ok=True)
            load(name, sources=[os.path.join(opdir, f) for f in files], build_directory=bdir, verbose=verbose,
                 is_python_module=False)
            shutil.copy(os.path.join(bdir, name + ".so"), target)
            shutil.rmtree(bdir, ignore_errors=True)
        built.append(target)
    built.append(build_refpy())
    return built


REFPY = os.path.join(OUT, "refpy")
# models/stylegan2/op JIT-builds CUDA at import and IS the kernel boundary; antialiased_sampling.py is the boundary too (the
# shim registers this repo's module under its name first, so the compiled file is never imported by the drop-in tests) but
# it is compiled so that tools/opbench.py --stn can TIME the reference's own MipmapWarp on the GPU next to the fused sampler
_REFPY_SKIP = (os.path.join("models", "stylegan2", "op"),)
_REFPY_UTILS = ("__init__.py", "distributed.py", "download.py", "annealing.py")


# A second, COMPLETE tree (models/stylegan2/op included) for the CPU arm of bench.py: with torch.utils.cpp_extension.load
# stubbed, the reference's ops take their own native CPU branches (op/upfirdn2d.py:146-149, op/fused_act.py:87-94), so
# `bench.py --impl reference` times the UNMODIFIED reference algorithm on the host cores (oracle/reference_step.py).
REFPY_CPU = os.path.join(OUT, "refpy_cpu")


def build_refpy(root=None, skip=None):
    """Byte-compile the reference's network code into oracle/_ref/refpy (sourceless .pyc tree); with `root`/`skip`
    given, into another tree (refpy_cpu: nothing skipped)."""
    import py_compile
    if root is None:
        build_refpy(REFPY_CPU, ())
        root, skip = REFPY, _REFPY_SKIP
    todo = []
    for base, _, files in os.walk(os.path.join(REF, "models")):
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(base, f), REF)
                if not any(rel.startswith(s) for s in skip):
                    todo.append(rel)
    todo += [os.path.join("utils", f) for f in _REFPY_UTILS if os.path.exists(os.path.join(REF, "utils", f))]
    for rel in todo:
        dst = os.path.join(root, rel + "c")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: tracebacks name the reference file, not a path inside this repo
        py_compile.compile(os.path.join(REF, rel), cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return root


def refpy_available():
    return os.path.exists(os.path.join(REFPY, "models", "__init__.pyc"))


def refpy_cpu_available():
    return os.path.exists(os.path.join(REFPY_CPU, "models", "stylegan2", "op", "__init__.pyc"))


def load_ref(name):
    """Import a prebuilt reference pybind module (upfirdn2d_ref / fused_ref) from oracle/_ref, or None."""
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_splat_ref():
    """ctypes handle of the reference splat kernel launcher, or None."""
    import ctypes
    path = os.path.join(OUT, "libsplat_ref.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.SplatForwardGpu.restype = None
    lib.SplatForwardGpu.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 5
    return lib


if __name__ == "__main__":
    for p in build(verbose="-v" in sys.argv):
        print(p)
