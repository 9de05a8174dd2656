"""Build the sm_100a shared library (and the CLI) in-tree with nvcc.

    python -m centrifuger_b200.build

The library has no torch / pybind dependency: it is a plain C-ABI shared object
(include/centrifuger_b200.h) linked against the CUDA runtime only.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcfrb200.so")
CLI = os.path.join(HERE, "centrifuger-b200")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "centrifuger_b200.h"))
    return deps


def build(force=False, verbose=False):
    deps = sources()
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB,
                                        os.path.join(CSRC, "cfr_api.cu"),
                                        os.path.join(CSRC, "cfr_format.cpp")]  # cudart is linked statically (nvcc default)
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    main_cpp = os.path.join(CSRC, "cfr_main.cpp")
    if os.path.exists(main_cpp) and (force or _stale(CLI, deps + [LIB])):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", CLI, main_cpp, "-L" + HERE, "-lcfrb200",
               "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"]  # host-only C++: no CUDA code in the CLI
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
