"""Build the C-ABI CUDA library in-tree: motionpriorcmax_b200/_lib/libcmax_b200.so (sm_100a only).

    python -m motionpriorcmax_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(OUT_DIR, "libcmax_b200.so")
SOURCES = ["api.cu", "lut_stage.cu", "event_stage.cu", "tile_stage.cu", "image_stage.cu", "voxel_stage.cu", "flow_stage.cu"]
HEADERS = [os.path.join(CSRC, "cmax_common.cuh"),
           os.path.join(os.path.dirname(PKG), "include", "cmax_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # IEEE arithmetic as the reference evaluates it op by op: no FMA contraction, no fast math
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.join(CSRC, "host_pack.cpp")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    # host-only translation unit (loader-side packer): g++ with OpenMP
    hobj = os.path.join(OUT_DIR, "host_pack.o")
    # -msse4.1: floorf / truncf inline as roundss instead of libm calls (every x86-64 CPU since 2008)
    hcmd = ["g++", "-O3", "-msse4.1", "-std=c++17", "-fPIC", "-fopenmp", "-c", os.path.join(CSRC, "host_pack.cpp"), "-o", hobj]
    if verbose:
        print(" ".join(hcmd))
    subprocess.run(hcmd, check=True)
    objs.append(hobj)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stdout.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "shared", "-lgomp"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
