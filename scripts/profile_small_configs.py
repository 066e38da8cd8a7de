"""Where the public fit of the small configurations (C1: n = 10k, C2: n = 100k) spends its time, per tensor path."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from neo_ls_svm_b200 import NeoLSSVM, _lib  # noqa: E402
from neo_ls_svm_b200.datasets import make_churn_rows, make_regression_rows  # noqa: E402

X1, y1 = make_regression_rows(12_000, 20, n_informative=10)
X2, y2 = make_churn_rows(115_000, 70, 20)
NeoLSSVM().fit(X1[:2000], y1[:2000])
for core in ("ozaki", "ozaki-dmma-sweep", "dmma"):
    ctx = _lib.context(0)
    ctx.set_gemm_core(core)
    for name, X, y in (("c1", X1[:10_000], y1[:10_000]), ("c2", X2[:100_000], y2[:100_000])):
        for rep in range(2):
            ctx.profile(True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m = NeoLSSVM().fit(X, y)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            prof = ctx.profile_read()
            ctx.profile(False)
        print(core, name, f"fit {t*1e3:.1f} ms", {k: round(v * 1e3, 1) for k, v in m.fit_phases_.items()},
              {k: round(v["ms"], 2) for k, v in prof.items()}, flush=True)
