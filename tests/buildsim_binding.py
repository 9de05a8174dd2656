"""ctypes handle to tests/hostsim/libbuildsim.so -- TEST INFRASTRUCTURE ONLY.

The index builder (centrifuger_b200/csrc/cfr_build.cu) is written as data-parallel passes over a
thin execution layer (cfr_build_backend.cuh).  Compiled with g++ -DCFR_HOSTSIM those passes run as
plain host loops, so the builder's output can be compared with the reference builder's files in a
container without a GPU.  The shipped library contains the sm_100a build only and has no CPU path.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "centrifuger_b200", "csrc")
SRC = os.path.join(CSRC, "cfr_build.cu")
LIB = os.path.join(ROOT, "tests", "hostsim", "libbuildsim.so")


def load():
    deps = [SRC, os.path.join(CSRC, "cfr_build_backend.cuh"), os.path.join(ROOT, "include", "centrifuger_b200_build.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-x", "c++", "-DCFR_HOSTSIM", "-shared", "-fPIC",
                               "-o", LIB, SRC, "-lpthread"])
    return C.CDLL(LIB)
