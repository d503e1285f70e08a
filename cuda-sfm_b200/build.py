"""Builds libsfmb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python cuda-sfm_b200/build.py [--force]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsfmb200.so")
SOURCES = ["api.cu", "hypgen.cu", "score.cu", "geometry.cu", "small.cu", "refit.cu", "bundle.cu", "chain.cu", "mg.cu", "hostmath.cu", "la_wrappers.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-Wno-deprecated-gpu-targets",
    "-DSFMB200_PDL",      # programmatic dependent launch between the kernels of the path (internal.cuh); -1.2 % per step
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_variant(name: str, defines: list[str], sources: list[str] | None = None) -> str:
    """Experimental build with extra -D flags into tools/proto/explibs/<name>.so (A/B runs on the GPU box: tools/lib_ab.py,
    tools/tri_ab.py).  Only `sources` (default: all) are recompiled with the flags; the rest reuse the product objects."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    build()
    outdir = os.path.join(ROOT, "tools", "proto", "explibs")
    objdir = os.path.join(HERE, "build", "exp_" + name)
    os.makedirs(outdir, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        if sources is None or src in sources:
            obj = os.path.join(objdir, src + ".o")
            r = subprocess.run([nvcc, *NVCC_FLAGS, *defines, "-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        else:
            obj = os.path.join(HERE, "build", src + ".o")
        objs.append(obj)
    out = os.path.join(outdir, name + ".so")
    r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "sfmb200.h"))
    if not force and _newer(LIB, srcs + hdrs):
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if not force and _newer(obj, [src] + hdrs):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
