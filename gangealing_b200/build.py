"""Ahead-of-time build of libgg_b200.so (sm_100a only) with plain nvcc -- no import-time JIT.

The reference JIT-compiles its extensions at import (models/stylegan2/op/upfirdn2d.py:9-16,
fused_act.py:10-17, utils/splat2d_cuda/functional.py:9-27); here the library is built once, in-tree,
and shipped as a single C-ABI shared object that is loaded with ctypes (see _lib.py).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libgg_b200.so")
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libgg_b200.so cannot be built")
    return nvcc


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path):
    h = hashlib.sha1()
    deps = [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    deps += sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE))
    for d in deps:
        with open(d, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libgg_b200.so.  Incremental (content-hashed)."""
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    objs, procs, relinked = [], [], False
    for src in sources():
        name = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJDIR, name + ".o")
        stamp_file = obj + ".sha1"
        stamp = _stamp(src)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT), src, stamp_file, stamp))
    for proc, src, stamp_file, stamp in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode()))
        with open(stamp_file, "w") as fh:
            fh.write(stamp)
        relinked = True
    if relinked or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s" % res.stdout.decode())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
