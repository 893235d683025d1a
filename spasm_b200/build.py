"""Build libspasm_b200.so in-tree (spasm_b200/lib/), for sm_100a only.

    python -m spasm_b200.build            # incremental
    python -m spasm_b200.build --force

Host C (csrc/host/*.c) is compiled with gcc, the CUDA side (csrc/gpu/*.cu) with
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo; nvcc cross-compiles without a GPU.
The result travels to the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import argparse
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
HOST_DIR = os.path.join(HERE, "csrc", "host")
GPU_DIR = os.path.join(HERE, "csrc", "gpu")
OBJ_DIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libspasm_b200.so")

GCC = "/usr/bin/gcc"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
                     "-ccbin", "/usr/bin/g++", f"-I{INCLUDE}", f"-I{GPU_DIR}"] + os.environ.get("SPASM_B200_NVCC_EXTRA", "").split()
GCC_FLAGS = ["-std=gnu11", "-O2", "-g", "-fPIC", "-Wall", "-Wextra", "-Wno-format-truncation", f"-I{INCLUDE}"]


def _newer(src: str, deps: list[str], out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(f"build failed: {cmd[0]} {cmd[-1]}")
    if r.stderr.strip() and os.environ.get("SPASM_B200_BUILD_VERBOSE"):
        sys.stderr.write(r.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = glob.glob(os.path.join(INCLUDE, "*.h")) + glob.glob(os.path.join(GPU_DIR, "*.cuh"))
    jobs = []
    objs = []
    for src in sorted(glob.glob(os.path.join(HOST_DIR, "*.c"))):
        obj = os.path.join(OBJ_DIR, "host_" + os.path.basename(src)[:-2] + ".o")
        objs.append(obj)
        if force or _newer(src, headers, obj):
            jobs.append([GCC] + GCC_FLAGS + ["-c", src, "-o", obj])
    for src in sorted(glob.glob(os.path.join(GPU_DIR, "*.cu"))):
        obj = os.path.join(OBJ_DIR, "gpu_" + os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, headers, obj):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj])
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_run, jobs))
    if jobs or not os.path.exists(LIB):
        _run([NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    if verbose:
        print(f"built {LIB} ({len(jobs)} objects recompiled)")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    build(force=a.force, verbose=True)
