"""Builds libsb2gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m starbeast2_b200.build [--force]

prepare.cu / reduce.cu / step.cu carry the reference's scalar arithmetic and are compiled with -fmad=false so that sums
and products associate exactly as in the Java; the pruning arithmetic (walk_common.cuh) uses explicit fma / rn intrinsics, so it
is the same in every translation unit.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsb2gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-ccbin", "/usr/bin/g++"]
SOURCES = {
    "prepare.cu": ["-fmad=false"],
    "reduce.cu": ["-fmad=false"],
    "treewalk.cu": [],
    "step.cu": ["-fmad=false"],
    "kstate.cu": ["-fmad=false"],
    "query.cu": ["-fmad=false"],
    "compress.cu": [],
    "engine.cu": [],
}
DEPS = ["sb2_device.cuh", "reduce_body.cuh", "prepare_body.cuh", "walk_common.cuh", os.path.join("..", "..", "include", "sb2gpu.h")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.abspath(__file__)]
    objs = []
    relink = force
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + deps):
            cmd = [NVCC] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            relink = True
    if relink or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
