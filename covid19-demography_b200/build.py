"""Build libseir_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python covid19-demography_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library links cudart statically and has no
Python / torch dependency: the boundary is the C ABI in include/seir_b200.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libseir_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "seir_capi.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "seir_kernels.cuh"),
                  os.path.join(ROOT, "include", "seir_b200.h"),
                  os.path.join(ROOT, "include", "seir_draws.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared", "--fmad=false", "-Xptxas", "-v"]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [NVCC] + FLAGS + os.environ.get("SEIR_NVCC_EXTRA", "").split() + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libseir_b200.so")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
