"""Builds lightzero_b200/_lib/liblzb200.so (the C-ABI library, include/lzb200.h) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
tree.cu is compiled with -fmad=false (the reference tree is built for baseline x86-64 and never
contracts a*b+c; bit-exact visit counts depend on it); the network kernels want FMA.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "liblzb200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math"]
if os.environ.get("LZ_UNIFORM_ISSUE"):      # experimental (round 2): elect.sync MMA issue in net_tc.cu / conv_tc.cu / ez.cu, see profiles/r01e_mma_probe.md
    COMMON.append("-DLZ_UNIFORM_ISSUE")
UNITS = [("tree.cu", ["-fmad=false"]), ("model.cu", []), ("net_tc.cu", []), ("conv_tc.cu", []), ("mlp.cu", []), ("ez.cu", []), ("search.cu", [])]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "lzb200.h"))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = ["nvcc"] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.check_call(cmd)
        objs.append(o)
    if force or _stale(LIB, objs):
        subprocess.check_call(["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
