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


def source_hash() -> str:
    """sha256 over every source the library is built from (csrc/*, include/*.h) and the compile flags: independent of
    mtimes, so a stale prebuilt .so is detectable (sfmb200_build_info() returns the hash it was built from)."""
    import hashlib

    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files += sorted(os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


LAST_BUILD: dict = {}


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
    objs.append(os.path.join(HERE, "build", "buildinfo.cu.o"))
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
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    src_hash = source_hash()
    stamp = os.path.join(objdir, "lib.hash")
    have = open(stamp).read().strip() if os.path.exists(stamp) else ""
    global LAST_BUILD
    if not force and os.path.exists(LIB) and have == src_hash:
        LAST_BUILD = {"compiled": 0, "src_hash": src_hash, "up_to_date": True}
        return LIB
    compiled = []
    info_src = os.path.join(objdir, "buildinfo.cu")
    with open(info_src, "w") as f:
        f.write('extern "C" const char* sfmb200_build_info(void) { return "src=%s nvcc_flags=%s"; }\n' % (src_hash, " ".join(NVCC_FLAGS)))

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if not force and _newer(obj, [src] + hdrs):
            return obj
        compiled.append(os.path.basename(src))
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    import time

    t0 = time.time()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs) + 1)) as ex:
        objs = list(ex.map(compile_one, srcs + [info_src]))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(src_hash)
    LAST_BUILD = {"compiled": len(compiled), "files": compiled, "src_hash": src_hash, "up_to_date": False, "seconds": round(time.time() - t0, 1)}
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(LAST_BUILD)
