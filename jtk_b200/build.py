"""Builds jtk_b200/libjtkgpu.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjtkgpu.so")
SOURCES = ["phmm_kernels.cu", "jtk_gpu_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "--expt-relaxed-constexpr"]


def sources():
    out = [os.path.join(CSRC, s) for s in SOURCES]
    for extra in sorted(os.listdir(CSRC)):
        if extra.endswith((".cu", ".cpp")) and extra not in SOURCES:
            out.append(os.path.join(CSRC, extra))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "jtk_gpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = LIB, extra=()) -> str:
    """`out` / `extra` build a tuning variant next to the product library (tools/prof.py, JTK_LIB_PATH)."""
    if out == LIB and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
