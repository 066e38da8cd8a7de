"""The host half of the INT8 (Ozaki) scheme, csrc/ozaki.cuh, checked on the CPU: digit extraction and reconstruction,
exponent rule, tile images, and whole dot products through the 28 exact level sums against long double.  The source
is compiled with nvcc (the `__host__ __device__` helpers are the ones the kernels use) and run on the host."""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ozaki_host_digit_scheme(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not available")
    exe = tmp_path / "ozaki_digits"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "neo_ls_svm_b200", "csrc"),
           "-o", str(exe), os.path.join(ROOT, "tests", "host_checks", "ozaki_digits.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok"), run.stdout + run.stderr
