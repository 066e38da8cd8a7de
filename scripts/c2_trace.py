import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_ls_svm_b200 import NeoLSSVM, _lib
from neo_ls_svm_b200.datasets import make_churn_rows
X2, y2 = make_churn_rows(115_000, 70, 20)
for core in ("ozaki", "ozaki-dmma-sweep"):
    ctx = _lib.context(0); ctx.set_gemm_core(core)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m = NeoLSSVM().fit(X2[:100_000], y2[:100_000])
        torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(core, f"{t*1e3:.1f} ms", {k: round(v*1e3,1) for k,v in m.fit_phases_.items() if k in ("feature_map_fit","solve")}, "D", m.primal_feature_map_.num_features if hasattr(m.primal_feature_map_,"num_features") else None, flush=True)
