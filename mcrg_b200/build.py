"""Builds mcrg_b200/libmcrg_b200.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcrg_b200.so")
SOURCES = ["capi.cu", "kernels.cu", "cluster.cu", "rgnn.cu", "util_kernels.cu", "comm.cu", "hostpack.cpp"]
HEADERS = ["bitops.cuh", "tile.cuh", "mcfast.cuh", "kernels.cuh", "capi_internal.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-t", "0", "-Xcompiler", "-fPIC",
              "-shared", "-ldl"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "mcrg_b200.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build_library(force=False, verbose=False):
    """nvcc cross-compiles without a GPU; a few seconds."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    build_library(force=True, verbose=True)
    print(LIB)
