"""Builds cm3_b200/csrc/libcm3env.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# (source, extra defines, object suffix): checkers.cu is compiled once per arithmetic type so the
# two halves of its template instantiations build in parallel
UNITS = [("api.cu", (), "api"), ("comm.cu", (), "comm"), ("particle.cu", (), "particle"),
         ("particle_pair.cu", (), "particle_pair"), ("particle_duo.cu", (), "particle_duo"),
         ("checkers.cu", ("CM3_CK_REAL=0",), "checkers_f32"),
         ("checkers.cu", ("CM3_CK_REAL=1",), "checkers_f64"),
         ("checkers.cu", ("CM3_CK_REAL=2",), "checkers_f32_i8"),
         ("checkers.cu", ("CM3_CK_REAL=3",), "checkers_f32_u2")]
SOURCES = ["api.cu", "comm.cu", "checkers.cu", "particle.cu", "particle_pair.cu", "particle_duo.cu"]
HEADERS = ["common.cuh", "params.cuh", "particle_math.cuh", os.path.join("..", "..", "include", "cm3env.h")]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def library_path():
    """The in-tree library; CM3ENV_LIBRARY selects an experimental build (build_library(defines=..., output=...),
    compared with tools/ab_r02.py)."""
    return os.environ.get("CM3ENV_LIBRARY") or os.path.join(CSRC, "libcm3env.so")


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


def build_library(force=False, verbose=False, defines=(), output=None):
    """Compile every CUDA source for sm_100a and link one shared library; returns its path.
    `defines` / `output` build an experimental variant next to the default library."""
    out = output or library_path()
    if not force and not defines and not is_stale():
        return out
    tag = os.path.splitext(os.path.basename(out))[0]
    # objects live outside the tree (132 MB of -lineinfo objects would travel with every gpurun
    # snapshot); only the linked library is kept in-tree
    objdir = os.path.join(os.environ.get("CM3_BUILD_DIR") or os.path.join(tempfile.gettempdir(), "cm3_b200_build"), tag)
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src, unit_defs, name in UNITS:
        obj = os.path.join(objdir, name + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
            ["-D%s" % d for d in tuple(defines) + tuple(unit_defs)] + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    subprocess.check_call([_nvcc()] + ARCH_FLAGS + ["-shared", "-o", out] + objs + ["-ldl"])
    return out
