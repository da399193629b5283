"""In-tree build of the hot path's native code for sm_100a.

  libmp_b200.so                      CUDA kernels + runtime behind the C ABI (include/*.h)
  millipyde.cpython-*.so             the CPython extension `millipyde` (drop-in API surface)

Plain nvcc / gcc invocations, no build system: `python -m millipyde_b200.build`.
The outputs live next to this file so they travel to the GPU box with the
repository snapshot (they are git-ignored).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmp_b200.so")
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
EXT = os.path.join(HERE, "millipyde" + EXT_SUFFIX)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
                     "-I", INCLUDE, "-I", CSRC]


def _newer(src_files, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in src_files)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_lib(force=False, verbose=False, ptxas_verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "kernels", "*.cuh")) + \
        glob.glob(os.path.join(INCLUDE, "*.h"))
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    objs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            extra = ["-Xptxas", "-v"] if ptxas_verbose and src.endswith(".cu") else []
            _run([NVCC] + NVCC_FLAGS + extra + ["-x", "cu", "-c", src, "-o", obj], verbose)
    if force or _newer(objs, LIB):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread"], verbose)
    return LIB


def build_ext(force=False, verbose=False):
    import numpy
    srcs = sorted(glob.glob(os.path.join(CSRC, "py", "*.c")))
    if not srcs:
        return None
    hdrs = glob.glob(os.path.join(CSRC, "py", "*.h")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    if force or _newer(srcs + hdrs + [LIB], EXT):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", "-Wno-unused-function",
               "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               "-I", INCLUDE, "-I", os.path.join(CSRC, "py"),
               "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include()] + srcs + \
              ["-o", EXT, "-L", HERE, "-lmp_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"]
        _run(cmd, verbose)
    return EXT


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    ext = build_ext(force, verbose)
    return lib, ext


if __name__ == "__main__":
    force = "--force" in sys.argv
    pv = "--ptxas-v" in sys.argv
    lib = build_lib(force, True, pv)
    ext = build_ext(force, True)
    print("built", lib, ext)
