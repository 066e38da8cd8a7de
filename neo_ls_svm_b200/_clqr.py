"""Coherent (non-crossing) linear quantile regression on the conformal calibration rows.

Mirror of /root/reference/src/neo_ls_svm/_coherent_linear_quantile_regressor.py.  The fit is a small
sparse LP (≤1440 rows) solved by HiGHS on the host — out of the GPU hot path's scope;
only its `predict` (an F×Q matmul per row) is folded into the GPU quantile epilogue
(`nls_quantile_epilogue`).  The LP is assembled with the reference's variable and constraint ordering
(:104-172) so that HiGHS follows the same path and returns the same vertex.
"""

from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.optimize import linprog
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.utils.validation import check_array, check_consistent_length, check_is_fitted, check_X_y


def coherent_linear_quantile_regression(X, y, *, quantiles, sample_weight=None, coherence_buffer: int = 3):
    """Minimise the pinball loss of ŷ⁽ʲ⁾ = Xβ⁽ʲ⁾ over all quantile ranks subject to Xβ⁽ʲ⁾ ≤ Xβ⁽ʲ⁺¹⁾.

    Variables, in order: β (Q·F, free), t = |β| (Q·F, ≥0), Δ⁺ (Q·n, ≥0), Δ⁻ (Q·n, ≥0) with
    Xβ⁽ʲ⁾ − y = Δ⁺⁽ʲ⁾ − Δ⁻⁽ʲ⁾.  `coherence_buffer` auxiliary ranks are inserted between consecutive
    requested ranks to strengthen monotonicity.  Returns (β for the requested ranks, β for all ranks).
    """
    n, F = X.shape
    dt = X.dtype
    n_req = len(quantiles)
    quantiles = np.interp(
        x=np.linspace(0, n_req - 1, (n_req - 1) * (1 + coherence_buffer) + 1), xp=np.arange(n_req), fp=quantiles
    ).astype(quantiles.dtype)
    Q = len(quantiles)
    assert np.array_equal(quantiles, np.sort(quantiles)), "Quantile ranks must be sorted."
    assert sample_weight is None or np.all(sample_weight >= 0), "Sample weights must be >= 0."
    sample_weight = np.ones(n, dtype=y.dtype) if sample_weight is None else sample_weight
    sample_weight /= np.sum(sample_weight)
    l1 = np.sqrt(np.finfo(y.dtype).eps) / (Q * F)  # tiny L1 penalty that only breaks ties
    cost = np.hstack([
        np.zeros(Q * F, dtype=y.dtype),
        l1 * np.ones(Q * F, dtype=y.dtype),
        np.kron((1 - quantiles) / Q, sample_weight),
        np.kron(quantiles / Q, sample_weight),
    ])
    eye_qn = sparse.eye(Q * n, dtype=dt)
    A_eq = sparse.hstack([
        sparse.kron(sparse.eye(Q, dtype=dt), X),
        sparse.csr_matrix((Q * n, Q * F), dtype=dt),
        -eye_qn,
        eye_qn,
    ])
    b_eq = np.tile(y, Q)
    eye_qf = sparse.eye(Q * F, dtype=dt)
    no_delta = sparse.csr_matrix((Q * F, 2 * Q * n), dtype=dt)
    no_beta = sparse.csr_matrix(((Q - 1) * n, 2 * Q * F), dtype=dt)
    eye_n = sparse.eye(n, dtype=dt)
    step_up = sparse.diags(diagonals=[1, -1], offsets=[0, 1], shape=(Q - 1, Q), dtype=dt)
    step_dn = sparse.diags(diagonals=[-1, 1], offsets=[0, 1], shape=(Q - 1, Q), dtype=dt)
    A_ub = sparse.vstack([
        sparse.hstack([eye_qf, -eye_qf, no_delta]),  # β ≤ t
        sparse.hstack([-eye_qf, -eye_qf, no_delta]),  # −β ≤ t
        sparse.hstack([no_beta, sparse.kron(step_up, eye_n), sparse.kron(step_dn, eye_n)]),  # monotone
    ])
    b_ub = np.zeros(A_ub.shape[0], dtype=dt)
    bounds = [(None, None)] * (Q * F) + [(0, None)] * (Q * F) + [(0, None)] * (2 * Q * n)
    sol = linprog(c=cost, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method="highs")
    beta_full = sol.x[: Q * F].astype(y.dtype).reshape(Q, F).T
    return beta_full[:, 0 :: (coherence_buffer + 1)], beta_full


class CoherentLinearQuantileRegressor(RegressorMixin, BaseEstimator):
    """Linear model regressing several quantiles coherently (non-crossing)."""

    def __init__(self, *, quantiles=(0.025, 0.5, 0.975), fit_intercept: bool = True, coherence_buffer: int = 3):
        self.quantiles = quantiles
        self.fit_intercept = fit_intercept
        self.coherence_buffer = coherence_buffer

    def _design(self, X):
        return np.hstack([X, np.ones((X.shape[0], 1), dtype=X.dtype)]) if self.fit_intercept else X

    def fit(self, X, y, *, sample_weight=None):
        X, y = check_X_y(X, y, dtype=(np.float64, np.float32), y_numeric=True)
        self.n_features_in_ = X.shape[1]
        self.y_dtype_ = X.dtype if np.issubdtype(y.dtype, np.integer) else y.dtype
        if np.issubdtype(y.dtype, np.datetime64) or np.issubdtype(y.dtype, np.timedelta64):
            X, y = X.astype(np.float64), y.astype(np.float64)
        y = y.astype(X.dtype)
        if sample_weight is not None:
            check_consistent_length(y, sample_weight)
            sample_weight = np.asarray(sample_weight).astype(y.dtype)
        self.β_, self.β_full_ = coherent_linear_quantile_regression(
            self._design(X), y, quantiles=np.asarray(self.quantiles).astype(y.dtype),
            sample_weight=sample_weight, coherence_buffer=self.coherence_buffer,
        )
        return self

    def predict(self, X):
        check_is_fitted(self)
        X = check_array(X, dtype=self.β_.dtype)
        yhat = self._design(X) @ self.β_
        return np.squeeze(yhat, axis=1 if yhat.shape[1] == 1 else ())

    def intercept_clip(self, X, y):
        """Range by which each quantile's intercept may move without crossing its neighbours (:257-272)."""
        check_is_fitted(self)
        X, y = check_X_y(X, y, dtype=self.β_.dtype, y_numeric=True)
        R = self._design(X) @ self.β_full_ - y[:, np.newaxis]
        clip = np.vstack([
            np.insert(np.max(R[:, :-1] - R[:, 1:], axis=0), 0, -np.inf),
            np.append(np.min(R[:, 1:] - R[:, :-1], axis=0), np.inf),
        ])
        clip[:, clip[0, :] >= clip[1, :]] = 0
        return clip[:, 0 :: (self.coherence_buffer + 1)]
