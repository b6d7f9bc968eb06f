"""Builds libcpmusic.so (all CUDA kernels + the C-ABI) in-tree for sm_100a with nvcc.

    python reinforcement-learning-in-music-generation_b200/build.py [--force]

Objects go to csrc/_build/ (git-ignored), the library next to this file so it travels to
the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcpmusic.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def _newer(a: str, deps) -> bool:
    return os.path.exists(a) and all(os.path.getmtime(a) >= os.path.getmtime(d) for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  [os.path.join(HERE, "..", "include", "cpmusic.h")])
    bdir = os.path.join(CSRC, "_build")
    os.makedirs(bdir, exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(cc, jobs):
                if verbose and log:
                    print(log)
    if jobs or force or not _newer(OUT, objs):
        # the CUDA runtime is linked dynamically (libcudart.so.12: the copy torch has already loaded, else the toolkit's)
        cmd = [NVCC, "-shared", "-cudart", "shared", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", OUT] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
