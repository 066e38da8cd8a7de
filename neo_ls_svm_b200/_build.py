"""Build recipe for libnls_b200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun)."""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnls_b200.so")
SOURCES = ["nls_b200.cu"]
HEADERS = ["ptx.cuh", "gemm_core.cuh", "ops.cuh", "small_kernels.cuh", "jacobi.cuh", "jacobi_wide.cuh", "binstats.cuh", "secular.h", "stedc_host.h", "stedc.cuh", "hetrd.cuh", "potrf.cuh", "ozaki.cuh", "eig_driver.inc", "../../include/nls_b200.h"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(os.path.normpath(f)) > t for f in deps if os.path.exists(os.path.normpath(f)))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a if it is missing or older than its sources."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-Xcompiler", "-fPIC", "--use_fast_math=false" if False else "-Xcompiler", "-O2",
        "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcusolver", "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
