"""Wall-clock phases of the public NeoLSSVM.fit at n = 4M (device synchronised at every phase boundary)."""
import os
import sys
import time
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import neo_ls_svm_b200._affine as aff  # noqa: E402
import neo_ls_svm_b200._neo_ls_svm as est  # noqa: E402
import neo_ls_svm_b200._primal as prim  # noqa: E402
import neo_ls_svm_b200._binstats as bst  # noqa: E402
import neo_ls_svm_b200._quantizer as qz  # noqa: E402
from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures  # noqa: E402
from neo_ls_svm_b200.datasets import fast_regression_rows  # noqa: E402

T = OrderedDict()


def wrap(mod, name, label=None):
    fn = getattr(mod, name)
    label = label or f"{mod.__name__.split('.')[-1]}.{name}"

    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.synchronize()
            T[label] = T.get(label, 0.0) + time.perf_counter() - t0

    setattr(mod, name, inner)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
X, y = fast_regression_rows(n, 64, 32)
mk = lambda: NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=1024), dual=False)  # noqa: E731
mk().fit(X[:50_000], y[:50_000])
for mod, name in [(est, "check_X_y"), (est, "unique_values"), (est, "train_test_split"), (aff, "_target_bins"), (aff, "_bin_location_spread"),
                  (aff, "_weighted_draw"), (aff, "nearest_neighbours"), (prim, "primal_fit"), (qz, "sample_bins_quantized_ecdf"),
                  (aff, "check_X_y")]:
    if hasattr(mod, name):
        wrap(mod, name)
wrap(est.NeoLSSVM, "_optimize_β̂_γ", "estimator._optimize")
wrap(aff.AffineSeparator, "fit", "AffineSeparator.fit")
for rep in range(2):
    T.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mk().fit(X, y)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(f"fit {total:.3f} s")
    for k, v in T.items():
        print(f"  {k:45s} {v * 1e3:8.1f} ms")
