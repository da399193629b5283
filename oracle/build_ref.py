"""CPU oracle, part C: build the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE ONLY.  The reference is a HIP C-extension whose own build
needs /opt/rocm-4.5.0/hip/bin/hipcc (setup.py:51-53) -- not runnable here.  Its
sources do compile for sm_100a as they lie, given one shim header that maps the
~35 HIP names they use onto the CUDA runtime.  This
script writes that shim into oracle/_ref/shim/hip/hip_runtime.h, compiles the
four .cpp files with nvcc and the eight .c files with gcc straight from
/root/reference/src, and links oracle/_ref/millipyde.cpython-*.so.  Nothing from
the reference is copied into the repository; oracle/_ref/ is git-ignored (but
travels to the GPU box with the snapshot).

The result needs a GPU to import (the module initialises devices at import), so
it is used only under gpurun: tests/test_reference_crosscheck.py compares the
product with it bit for bit on the reference's own layouts, and
tests/golden/make_golden.py records its outputs as fixtures.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("MILLIPYDE_REFERENCE", "/root/reference")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SHIM = r"""// HIP -> CUDA name shim for building jasbury1/millipyde unmodified (test infrastructure).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#define hipError_t cudaError_t
#define hipSuccess cudaSuccess
#define hipStream_t cudaStream_t
#define hipDeviceProp_t cudaDeviceProp
#define hipDeviceptr_t void *
#define hipMemcpyHostToDevice cudaMemcpyHostToDevice
#define hipMemcpyDeviceToHost cudaMemcpyDeviceToHost
#define hipMemcpyDeviceToDevice cudaMemcpyDeviceToDevice
#define hipGetErrorString cudaGetErrorString
#define hipSetDevice cudaSetDevice
#define hipMalloc cudaMalloc
#define hipFree cudaFree
#define hipMemcpy cudaMemcpy
#define hipMemcpyPeerAsync cudaMemcpyPeerAsync
#define hipStreamCreate cudaStreamCreate
#define hipStreamSynchronize cudaStreamSynchronize
#define hipDeviceSynchronize cudaDeviceSynchronize
#define hipDeviceReset cudaDeviceReset
#define hipGetDeviceCount cudaGetDeviceCount
#define hipGetDeviceProperties cudaGetDeviceProperties
#define hipDeviceCanAccessPeer cudaDeviceCanAccessPeer
#define hipDeviceEnablePeerAccess cudaDeviceEnablePeerAccess
#define hipMemcpyDtoD(d, s, n) cudaMemcpy((d), (s), (n), cudaMemcpyDeviceToDevice)
#define HIP_SYMBOL(x) x
#define hipMemcpyToSymbol cudaMemcpyToSymbol
#define hipThreadIdx_x threadIdx.x
#define hipThreadIdx_y threadIdx.y
#define hipBlockIdx_x blockIdx.x
#define hipBlockIdx_y blockIdx.y
#define hipBlockDim_x blockDim.x
#define hipBlockDim_y blockDim.y
#define hipLaunchKernelGGL(k, g, b, sh, st, ...) k<<<(g), (b), (sh), (st)>>>(__VA_ARGS__)
"""


def so_path() -> str:
    return os.path.join(OUT, "millipyde" + sysconfig.get_config_var("EXT_SUFFIX"))


def available() -> bool:
    return os.path.exists(so_path())


def build(verbose: bool = True, force: bool = False) -> str:
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src}: reference sources not present (expected on the GPU box)")
    import numpy
    os.makedirs(os.path.join(OUT, "shim", "hip"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    with open(os.path.join(OUT, "shim", "hip", "hip_runtime.h"), "w") as f:
        f.write(SHIM)
    target = so_path()
    if available() and not force:
        return target
    inc = ["-I", os.path.join(OUT, "shim"), "-I", os.path.join(src, "include"),
           "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include()]
    objs = []

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL,
                              stderr=None if verbose else subprocess.DEVNULL)

    for cpp in sorted(glob.glob(os.path.join(src, "*.cpp"))):
        obj = os.path.join(OUT, "obj", os.path.basename(cpp) + ".o")
        run([NVCC, "-x", "cu", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-w",
             "-Xcompiler", "-fPIC", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc + ["-c", cpp, "-o", obj])
        objs.append(obj)
    for c in sorted(glob.glob(os.path.join(src, "*.c"))):
        obj = os.path.join(OUT, "obj", os.path.basename(c) + ".o")
        run(["gcc", "-fPIC", "-O2", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc + ["-c", c, "-o", obj])
        objs.append(obj)
    run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + objs + ["-lpthread"])
    return target


# ---------------------------------------------------------------------------------------------
# The boundary proven by construction: the reference's seven CPython type files, compiled unmodified
# from /root/reference/src against the reference's own headers, linked against libmp_b200.so INSTEAD
# of its four HIP translation units (setup.py:48-66 source list minus millipyde_image.cpp,
# millipyde_objects.cpp, millipyde_devices.cpp, millipyde_workers.cpp; millipyde.c's mperr_str /
# random_* come from the library too).  If a symbol the reference's host code calls were missing or
# had another signature, this would not link or not run.
HYBRID_SOURCES = ["device.c", "gpuarray.c", "gpugenerator.c", "gpuimage.c", "gpuoperation.c", "gpupipeline.c",
                  "millipyde_module.c"]


def hybrid_so_path() -> str:
    return os.path.join(OUT, "hybrid", "millipyde" + sysconfig.get_config_var("EXT_SUFFIX"))


def hybrid_available() -> bool:
    return os.path.exists(hybrid_so_path())


def build_hybrid(verbose: bool = True, force: bool = False) -> str:
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src}: reference sources not present (expected on the GPU box)")
    import numpy
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "millipyde_b200")
    lib = os.path.join(libdir, "libmp_b200.so")
    if not os.path.exists(lib):
        raise FileNotFoundError(f"{lib}: build the product first (python -m millipyde_b200.build)")
    target = hybrid_so_path()
    os.makedirs(os.path.dirname(target), exist_ok=True)
    if hybrid_available() and not force and os.path.getmtime(target) >= os.path.getmtime(lib):
        return target
    inc = ["-I", os.path.join(src, "include"), "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include()]
    objs = []
    for name in HYBRID_SOURCES:
        obj = os.path.join(OUT, "hybrid", name + ".o")
        cmd = ["gcc", "-fPIC", "-O2", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc + \
              ["-c", os.path.join(src, name), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    # --no-undefined would also flag the CPython symbols an extension leaves to the interpreter, so the
    # check is explicit: every undefined non-Python symbol must be one libmp_b200.so (or libc) defines
    cmd = ["gcc", "-shared", "-o", target] + objs + ["-L", libdir, "-lmp_b200",
                                                     "-Wl,-rpath,$ORIGIN/../../../millipyde_b200", "-lpthread"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return target


def load_hybrid():
    """Import the reference's host code bound to libmp_b200.so (GPU required)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("millipyde", hybrid_so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import the shim-built reference as module `millipyde_reference` (GPU required)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("millipyde", so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    if "--hybrid" in sys.argv:
        print(build_hybrid(verbose=True, force="--force" in sys.argv))
    else:
        print(build(verbose=True, force="--force" in sys.argv))
