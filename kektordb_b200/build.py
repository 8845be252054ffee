"""Builds kektordb_b200/libkektordb_gpu.so (sm_100a only) with nvcc, in-tree.

The built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkektordb_gpu.so")
SOURCES = ["api.cu", "search.cu", "search_k0.cu", "search_k1.cu", "search_k2.cu", "search_k3.cu", "flat.cu", "flat_tc.cu", "build.cu", "arena.cu", "shard.cu", "batcher.cpp", "graphfile.cpp", "refresher.cpp"]
HEADERS = ["kdb_internal.cuh", "handle.h", "searcher.cuh", "search_inst.cuh", os.path.join("..", "..", "include", "kektordb_gpu.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    common = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-ccbin", host_cxx,
        "-Xcompiler", "-fPIC,-O2,-Wall",
        "--fmad=false",  # every FMA in the kernels is an explicit __fmaf_rn
        "-Xptxas", "-v" if verbose else "-O3",
    ]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:  # one nvcc per translation unit, in parallel
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = common + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, failed = [], False
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
        objs.append(obj)
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", host_cxx, "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
