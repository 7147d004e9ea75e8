"""Builds libpfgpu.so (the CUDA C-ABI library) in-tree with nvcc for sm_100a.

`python -m ploidyfrost_b200.build [--force] [-v]` or build_library().  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpfgpu.so")
SOURCES = ["pf_api.cu", "pf_kmc.cu", "pf_align.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    inc = os.path.join(ROOT, "include")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(inc, f) for f in ("pf_gpu.h", "pf_types.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # one object per translation unit, compiled side by side (objects under build/, git-ignored), then one link
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    hdr_t = max(os.path.getmtime(os.path.join(d, f)) for d in (CSRC, inc) for f in os.listdir(d) if f.endswith((".cuh", ".h", ".hpp")))
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            continue
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-I", inc, "-c", src, "-o", obj]
        jobs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    failed = False
    for s, pr in jobs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            sys.stderr.write(out + err)
            failed = True
        elif verbose:
            sys.stderr.write(err)
    if failed:
        raise RuntimeError("nvcc failed building libpfgpu.so")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libpfgpu.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
