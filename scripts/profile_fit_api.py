"""Where the public-API NeoLSSVM.fit spends its time at n = 4M (host + device), via cProfile."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures  # noqa: E402
from neo_ls_svm_b200.datasets import fast_regression_rows  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
X, y = fast_regression_rows(n, 64, 32)
mk = lambda: NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=1024), dual=False)  # noqa: E731
mk().fit(X[:50_000], y[:50_000])
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
mk().fit(X, y)
torch.cuda.synchronize()
pr.disable()
print("fit seconds", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
