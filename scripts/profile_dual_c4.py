"""Stage split of the dual fit at C4 (n = 16,384, d = 32) from the library's per-kind CUDA-event profile."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from neo_ls_svm_b200 import NeoLSSVM, _lib  # noqa: E402
from neo_ls_svm_b200.datasets import make_regression_rows  # noqa: E402

X, y = make_regression_rows(16_384, 32, n_informative=16)
NeoLSSVM(dual=True).fit(X[:2000], y[:2000])
ctx = _lib.context()
for rep in range(2):
    ctx.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = NeoLSSVM(dual=True).fit(X, y)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile(False)
    print(f"fit {dt:.3f} s; kinds (ms):", {k: round(v["ms"], 1) for k, v in prof.items()}, "launches", {k: v["launches"] for k, v in prof.items()})
