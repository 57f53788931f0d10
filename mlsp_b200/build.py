"""Builds libmlsp_b200.so (the C-ABI library of include/mlsp_b200.h) in-tree with nvcc for sm_100a.

    python -m mlsp_b200.build [--force] [--verbose]

No torch involvement: the library exports plain `extern "C"` symbols and is loaded with ctypes
(mlsp_b200/_lib.py).  The .so lands in mlsp_b200/lib/ (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libmlsp_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(f for f in os.listdir(SRC) if f.endswith(".cu"))


def _newest_header() -> float:
    t = os.path.getmtime(os.path.join(INCLUDE, "mlsp_b200.h"))
    for f in os.listdir(SRC):
        if f.endswith((".cuh", ".h")):
            t = max(t, os.path.getmtime(os.path.join(SRC, f)))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for f in sources():
        src = os.path.join(SRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stdout + r.stderr)

    if jobs:
        with ThreadPoolExecutor(min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
