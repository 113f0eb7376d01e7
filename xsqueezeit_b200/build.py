"""Builds xsqueezeit_b200/libxsi_b200.so in-tree: nvcc for sm_100a (hand-written kernels + C ABI)
and g++ for the host container layer.  Cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libxsi_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["xsi_b200.cu", "xsi_container.cpp", "host_narrow.cpp"]
DEPS = SOURCES + ["common.cuh", "encode_kernels.cuh", "decode_kernels.cuh", "host_util.hpp", "host_narrow.hpp",
                  os.path.join("..", "..", "include", "xsi_b200.h")]


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s and %s is missing or stale" % (NVCC, SO))
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC,-O2", "-shared", "-diag-suppress", "550", "-o", SO] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lpthread"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
