"""A/B of the public fit at n = 4M: device target binning on / off, 4 alternating fits each."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures
from neo_ls_svm_b200.datasets import fast_regression_rows
X, y = fast_regression_rows(4_000_000, 64, 32)
mk = lambda: NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=1024), dual=False)
mk().fit(X[:50_000], y[:50_000]); mk().fit(X, y)
res = {"0": [], "1": []}
for rep in range(8):
    flag = str(rep & 1)
    os.environ["NLS_HOST_BINS"] = flag
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = mk().fit(X, y)
    torch.cuda.synchronize(); res[flag].append((time.perf_counter() - t0, m.fit_phases_["feature_map_fit"], m.fit_phases_["validation"]))
for flag, v in res.items():
    print("host bins" if flag == "1" else "device bins", "fit min %.3f s" % min(a for a, _, _ in v), "feature_map_fit min %.3f" % min(b for _, b, _ in v), "validation min %.3f" % min(c for _, _, c in v), [round(a, 3) for a, _, _ in v])
