"""Seeded synthetic workloads for the configurations named in BASELINE.json / SURVEY.md §8d.

There is no network on the build or GPU boxes, so every test and benchmark input is generated from
these recipes (all `random_state=0`; estimator seeds stay at their default 42).
"""

from __future__ import annotations

import numpy as np


def make_regression_rows(n: int, d: int, n_informative: int | None = None, noise: float = 1.0, seed: int = 0):
    """sklearn `make_regression` recipe used for C1/C3/C4/C5."""
    from sklearn.datasets import make_regression

    X, y = make_regression(
        n_samples=n,
        n_features=d,
        n_informative=n_informative if n_informative is not None else min(10, d),
        noise=noise,
        random_state=seed,
    )
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def make_churn_rows(n: int, d: int = 70, n_informative: int = 20, flip_y: float = 0.05, seed: int = 0):
    """Churn-shaped binary classification recipe used for C2 (~14% positives, 5% label noise)."""
    from sklearn.datasets import make_classification

    X, y = make_classification(
        n_samples=n,
        n_features=d,
        n_informative=n_informative,
        weights=[0.86, 0.14],
        flip_y=flip_y,
        random_state=seed,
    )
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


BLOCK_ROWS = 1 << 18


def fast_regression_rows(n: int, d: int, n_informative: int, noise: float = 1.0, seed: int = 0,
                         row_begin: int = 0, row_end: int | None = None):
    """Streaming equivalent of `make_regression` for multi-million-row benchmark inputs.

    Same distribution (standard-normal X, sparse linear ground truth with coefficients 100*U(0,1),
    Gaussian noise).  Rows are generated in blocks of 2^18, each block from its own
    `numpy.random.Generator` seeded by (seed, block), so any rank can materialise exactly its own row
    range [row_begin, row_end) of the same global dataset without generating the rest.
    """
    row_end = n if row_end is None else row_end
    coef = np.zeros(d)
    coef[:n_informative] = 100.0 * np.random.default_rng([seed, 1 << 30]).random(n_informative)
    X = np.empty((row_end - row_begin, d), dtype=np.float64)
    y = np.empty(row_end - row_begin, dtype=np.float64)
    for blk in range(row_begin // BLOCK_ROWS, (row_end + BLOCK_ROWS - 1) // BLOCK_ROWS):
        b0, b1 = blk * BLOCK_ROWS, min(n, (blk + 1) * BLOCK_ROWS)
        rng = np.random.default_rng([seed, blk])
        Xb = rng.standard_normal((b1 - b0, d))
        yb = Xb @ coef + noise * rng.standard_normal(b1 - b0)
        lo, hi = max(b0, row_begin), min(b1, row_end)
        X[lo - row_begin : hi - row_begin] = Xb[lo - b0 : hi - b0]
        y[lo - row_begin : hi - row_begin] = yb[lo - b0 : hi - b0]
    return X, y


CASES = {
    # name: (kind, kwargs, estimator kwargs)
    "reg_small": ("regression", dict(n=1200, d=6, n_informative=4, noise=50.0), dict(num_features=256, dual=False)),
    "clf_small": ("churn", dict(n=1500, d=10, n_informative=5, flip_y=0.2), dict(num_features=256, dual=False)),
    "c1": ("regression", dict(n=10_000, d=20, n_informative=10), dict()),
    "c2_small": ("churn", dict(n=6000, d=70, n_informative=20), dict()),
    "c3_small": ("regression", dict(n=5000, d=64, n_informative=32), dict(num_features=1024, dual=False)),
    "dual_reg": ("regression", dict(n=600, d=6, n_informative=4), dict(dual=True)),
    "dual_clf": ("churn", dict(n=500, d=10, n_informative=5), dict(dual=True)),
    # Full-size C2: the largest configuration the reference itself can run here (32 s).  Its fixture
    # stores the n-vectors on the 4096 rows of `golden_row_subset` only, which keeps the .npz small.
    "c2_full": ("churn", dict(n=100_000, d=70, n_informative=20), dict()),
}

GOLDEN_ROW_SUBSET = 4096


def golden_row_subset(n: int) -> np.ndarray:
    """Rows on which large fixtures keep their per-row vectors (all rows when n is small)."""
    if n <= 20_000:
        return np.arange(n)
    return np.linspace(0, n - 1, GOLDEN_ROW_SUBSET).astype(np.int64)


def load_case(name: str, n_test: int = 400):
    """Return (X_train, y_train, sample_weight, X_test, estimator_kwargs) for a golden case."""
    kind, kw, est = CASES[name]
    kw = dict(kw)
    n = kw.pop("n")
    if kind == "regression":
        X, y = make_regression_rows(n + n_test, **kw)
    else:
        X, y = make_churn_rows(n + n_test, **kw)
    rng = np.random.default_rng(1234)
    # Non-uniform weights for the small classification cases exercise the weighted code paths.
    sw = rng.uniform(0.5, 1.5, size=n) if name in ("clf_small", "dual_clf") else None
    return X[:n], y[:n], sw, X[n:], dict(est)
