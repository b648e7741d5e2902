"""Builds cm3_b200/csrc/libcm3env.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "checkers.cu", "particle.cu"]
HEADERS = ["common.cuh", "params.cuh", os.path.join("..", "..", "include", "cm3env.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def library_path():
    return os.path.join(CSRC, "libcm3env.so")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.isfile(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    lib = library_path()
    if not os.path.isfile(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    """Compile every CUDA source into one shared library; returns its path."""
    if not force and not is_stale():
        return library_path()
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", library_path()] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return library_path()
