"""Public NeoLSSVM.fit at n = 4M with rows sharded over the GPUs selected by NLS_DEVICES (single process)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures
from neo_ls_svm_b200.datasets import fast_regression_rows
X, y = fast_regression_rows(4_000_000, 64, 32)
mk = lambda: NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=1024), dual=False)
mk().fit(X[:50_000], y[:50_000])
out = []
for rep in range(3):
    for d in range(torch.cuda.device_count()): torch.cuda.synchronize(d)
    t0 = time.perf_counter(); m = mk().fit(X, y)
    for d in range(torch.cuda.device_count()): torch.cuda.synchronize(d)
    out.append(time.perf_counter() - t0)
print("NLS_DEVICES =", os.environ.get("NLS_DEVICES"), "fit seconds", [round(t, 3) for t in out], "gamma index", int(np.argmin(np.abs(m.γs_ - m.γ_))),
      {k: round(v, 3) for k, v in m.fit_phases_.items() if k in ("feature_map_fit", "solve", "calibration_split")}, flush=True)
