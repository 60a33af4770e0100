"""Build ``libpbp_engine.so`` in-tree with nvcc for sm_100a (no GPU needed to compile)."""

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libpbp_engine.so")
SOURCES = ["pbp_engine.cu"]
HEADERS = ["pbp_kernels.cuh", "pbp_host_tables.h", "pbp_surface.cuh", "pbp_gjk.cuh", "pbp_model.cuh", "pbp_ik.cuh", "pbp_supmap.h", "pbp_knn.cuh", "pbp_witness.cuh", "pbp_jit.h", os.path.join("..", "..", "include", "pbp_engine.h")]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # only explicit fmaf() fuses: device FP32 results match the host build used by the CPU tests
    "-Xcompiler",
    "-fPIC",
    "-shared",
    "-ldl",  # NVRTC (model-specialised broadphase, pbp_jit.h) and nothing else is dlopen'ed at run time
]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_engine(force=False, verbose=False):
    """Compile the CUDA engine. Returns the path of the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH] + [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    build_engine(force="--force" in sys.argv, verbose=True)
