"""Builds deltaq_b200/libdeltaq_cuda.so in-tree with nvcc for sm_100a (B200).

    python -m deltaq_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is git-ignored but travels to the GPU box with the
gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdeltaq_cuda.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-pthread,-Wall,-Wno-unused-function",
    "-shared", "-cudart", "static",
]


def sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    out.append(os.path.join(INCLUDE, "deltaq_cuda.h"))
    return out


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libdeltaq_cuda cannot be built (there is no CPU fallback)")


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in sources()):
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I", INCLUDE, "-o", OUT, os.path.join(CSRC, "deltaq_cuda.cu")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
