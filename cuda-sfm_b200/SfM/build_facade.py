"""Compiles the facade demo (reference driver lines + facade headers) with plain
g++ against libsfmb200.so: the "does the drop-in still compile" check."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(HERE, "facade_demo")


def build() -> str:
    src = os.path.join(HERE, "facade_demo.cpp")
    deps = [src] + [os.path.join(HERE, f) for f in ("sfm.h", "kernels.h", "svd.h", "common.h")] + [os.path.join(PKG, "libsfmb200.so")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wno-unused-function", f"-I{cuda}/include", f"-I{HERE}", "-o", OUT, src,
           f"-L{PKG}", "-lsfmb200", f"-L{cuda}/lib64", "-lcudart", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{cuda}/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("facade build failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build())
