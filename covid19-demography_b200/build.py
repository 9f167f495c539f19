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
DEPS = SOURCES + [os.path.join(HERE, "csrc", "seir_kernels.cuh"), os.path.join(HERE, "csrc", "seir_prologue_host.cuh"), os.path.join(HERE, "csrc", "seir_tsv_host.cuh"),
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
    """Compile to a temporary file and rename it into place under a file lock: bench.py --gpus N, the sharded and the
    ensemble runs start one process per GPU, and each of them lands here."""
    if not force and not is_stale():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():   # another process built it while this one waited
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [NVCC] + FLAGS + os.environ.get("SEIR_NVCC_EXTRA", "").split() + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libseir_b200.so")
            os.replace(tmp, LIB)
            with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:   # (git-ignored: registers / spills of this build)
                f.write(res.stdout)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
