"""Where NeoLSSVM().fit spends its time on C2 (churn-shaped binary classification, n = 100k, d = 70), via cProfile."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from neo_ls_svm_b200 import NeoLSSVM  # noqa: E402
from neo_ls_svm_b200.datasets import make_churn_rows  # noqa: E402

X, y = make_churn_rows(115_000, 70, 20)
Xtr, ytr, Xte = X[:100_000], y[:100_000], X[100_000:]
NeoLSSVM().fit(Xtr[:3000], ytr[:3000])
NeoLSSVM().fit(Xtr, ytr)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
m = NeoLSSVM().fit(Xtr, ytr)
torch.cuda.synchronize()
pr.disable()
print("fit seconds", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(38)
t0 = time.perf_counter(); m.predict_proba(Xte); print("predict_proba 15k s", time.perf_counter() - t0)
