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
    deps.append(os.path.join(ROOT, "include", "centrifuger_b200_build.h"))
    return deps


def _compile_objects(units, verbose):
    """nvcc -c every translation unit (in parallel), return the object paths."""
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + ["--extended-lambda", "-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        return list(ex.map(one, units))


def build(force=False, verbose=False):
    deps = sources()
    if force or _stale(LIB, deps):
        # classification path (cfr_api.cu), index builder (cfr_build.cu), file grammar (cfr_format.cpp):
        # one shared object, cudart linked statically (nvcc default)
        objs = _compile_objects([os.path.join(CSRC, "cfr_api.cu"), os.path.join(CSRC, "cfr_build.cu"),
                                 os.path.join(CSRC, "cfr_format.cpp")], verbose)
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    main_cpp = os.path.join(CSRC, "cfr_main.cpp")
    if os.path.exists(main_cpp) and (force or _stale(CLI, deps + [LIB])):
        cmd = ["g++", "-std=c++23", "-O2", "-Wall", "-o", CLI, main_cpp, "-L" + HERE, "-lcfrb200",  # (string::resize_and_overwrite)
               "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"]  # host-only C++: no CUDA code in the CLI
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    quant_cpp = os.path.join(CSRC, "cfr_quant_main.cpp")
    quant_cli = os.path.join(HERE, "centrifuger-b200-quant")
    if os.path.exists(quant_cpp) and (force or _stale(quant_cli, deps)):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", quant_cli, quant_cpp, os.path.join(CSRC, "cfr_format.cpp"), "-lz"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)  # host-only drop-in for centrifuger-quant (TSV in, report out)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
