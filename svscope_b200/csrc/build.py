"""Builds svscope_b200/_C/libsvscope_b200.so in-tree with nvcc for sm_100a.

    python -m svscope_b200.csrc.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT_DIR = os.path.join(PKG, "_C")
LIB = os.path.join(OUT_DIR, "libsvscope_b200.so")
OBJ_DIR = os.path.join(HERE, "_build")

CU_SOURCES = ["api.cu", "poa_kernels.cu", "poa_window.cu", "msa_features.cu", "em.cu", "myers.cu", "misscore.cu"]
CXX_SOURCES = []
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall", "--expt-relaxed-constexpr"]


def _newer(src, dst):
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "svscope_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs = []
    changed = False
    for src in CU_SOURCES + CXX_SOURCES:
        sp = os.path.join(HERE, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if force or _newer(sp, obj) or hdr_time > os.path.getmtime(obj):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            changed = True
    if changed or not os.path.exists(LIB):
        cmd = ["nvcc", "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB] + objs + ["-lcudart", "-lpthread"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
