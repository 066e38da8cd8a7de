"""Weighted quantiles (host pre-pass utility).

Behavioural mirror of the reference's `weighted_quantile`
(/root/reference/src/neo_ls_svm/_weighted_quantile.py:35-77): the q-th weighted quantile is the mean
of two linear interpolations of the sorted values, one against the cumulative weight *before* and
one *after* each sample.  Out of the GPU hot path's scope, kept in NumPy so that the
learned shift/scale match the reference.
"""

from __future__ import annotations

import numpy as np

try:  # numba is what the reference uses for the row-parallel interpolation; optional here.
    import numba

    @numba.jit(nopython=True, nogil=True, parallel=True, fastmath=True, cache=False)
    def _interp_rows(q, p, a):  # pragma: no cover - compiled
        res = np.empty((a.shape[0], len(q)), dtype=a.dtype)
        for r in numba.prange(a.shape[0]):
            res[r, :] = np.interp(q, p[r, :], a[r, :])
        return res

except Exception:  # noqa: BLE001

    def _interp_rows(q, p, a):
        res = np.empty((a.shape[0], len(q)), dtype=a.dtype)
        for r in range(a.shape[0]):
            res[r, :] = np.interp(q, p[r, :], a[r, :])
        return res


def weighted_quantile(a: np.ndarray, w: np.ndarray, q, axis: int | None = None) -> np.ndarray:
    """Weighted q-th quantile(s) of `a` with non-negative weights `w` along `axis` (or flattened)."""
    assert a.ndim == w.ndim, "Array and weights must have the same number of dimensions"
    assert axis is None or (0 <= axis < a.ndim), "Axis must be one of the array's dimensions"
    assert np.all(w >= 0), "Weights must be nonnegative"
    a = np.ascontiguousarray(a)
    w = np.broadcast_to(np.ascontiguousarray(w), a.shape)
    q = np.ravel(np.asarray([q])).astype(a.dtype)
    if axis is None:
        flat_a, flat_w = np.ravel(a), np.ravel(w)
        order = np.argsort(flat_a)
        flat_a, flat_w = flat_a[order], flat_w[order]
        cum = np.cumsum(flat_w)
        before, after = (cum - flat_w) / cum[-1], cum / cum[-1]
        return (0.5 * np.interp(q, before, flat_a) + 0.5 * np.interp(q, after, flat_a)).astype(flat_a.dtype)
    a, w = np.moveaxis(a, axis, -1), np.moveaxis(w, axis, -1)
    lead_shape = a.shape[:-1]
    a2, w2 = np.reshape(a, [-1, a.shape[-1]]), np.reshape(w, [-1, w.shape[-1]])
    order = np.argsort(a2, axis=1)
    a2, w2 = np.take_along_axis(a2, order, axis=1), np.take_along_axis(w2, order, axis=1)
    cum = np.cumsum(w2, axis=1)
    total = cum[:, [-1]].copy()
    before, after = (cum - w2) / total, cum / total
    res = (_interp_rows(q, before, a2) + _interp_rows(q, after, a2)) / 2
    res = np.reshape(res, lead_shape + (len(q),))
    return np.moveaxis(res, -1, axis)
