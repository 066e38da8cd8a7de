"""ECDF quantisation of a numeric vector into variable-width bins (host pre-pass utility).

Behavioural mirror of /root/reference/src/neo_ls_svm/_quantizer.py: bins are grown greedily from both
ends of the empirical CDF; a bin is closed as soon as a straight line through it would deviate from the
ECDF by more than `max_bin_error` samples, or it holds more than `max_bin_size` samples
(`_next_knot` :18-44, `_prev_knot` :47-73, `hist_quantized_ecdf` :104-177).  The target's bin index
drives the supervised normaliser / separator.  Out of the GPU hot path's scope.
"""

from __future__ import annotations

import os
from typing import Any

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils.validation import check_array


def _grow_bin_py(x, y, knot, max_err, max_size, forward):
    """Grow one bin from `knot` (forward: to the right; else to the left).

    Returns (stop_knot, samples_in_bin).  x: knot positions with -inf/+inf sentinels; y: cumulative
    counts with 0 / int-max sentinels.
    """
    lo_slope, hi_slope = 0.0, np.inf
    stop, count = knot, 0
    if forward:
        for stop in range(knot + 1, len(x)):
            count = int(y[stop - 1] - y[knot - 1] if knot > 0 else y[stop - 1])
            if count > max_size:
                break
            if stop == knot + 1:
                continue
            dx, dy = x[stop - 1] - x[knot], y[stop - 1] - y[knot]
            hi_slope = min(hi_slope, (dy + max_err) / dx)
            lo_slope = max(lo_slope, (dy - max_err) / dx)
            slope = dy / dx
            if not (lo_slope <= slope <= hi_slope):
                break
    else:
        for stop in range(knot - 1, -1, -1):
            count = int(y[knot - 1] - y[stop - 1] if stop > 0 else y[knot - 1])
            if count > max_size:
                break
            if knot == stop + 1:
                continue
            dx, dy = x[knot - 1] - x[stop], y[knot - 1] - y[stop]
            hi_slope = min(hi_slope, (dy + max_err) / dx)
            lo_slope = max(lo_slope, (dy - max_err) / dx)
            slope = dy / dx
            if not (lo_slope <= slope <= hi_slope):
                break
    return stop, count


try:
    import numba

    _grow_bin = numba.jit(nopython=True, nogil=True, fastmath=True, cache=False)(_grow_bin_py)
except Exception:  # noqa: BLE001
    _grow_bin = _grow_bin_py


# Above this many samples the distinct values of a vector are found with a device sort (same values, same
# inverse indices as np.unique, whose host argsort costs 0.3 s at n = 4M); the estimator needs a GPU anyway.
MIN_SAMPLES_FOR_DEVICE_UNIQUE = 1 << 20


def unique_values(x: np.ndarray, *, return_inverse: bool = False, return_counts: bool = False):
    """np.unique(x, return_inverse=..., return_counts=...) for a 1-D numeric vector, on the GPU when it is large."""
    x = np.asarray(x)
    if x.ndim == 1 and x.size >= MIN_SAMPLES_FOR_DEVICE_UNIQUE and x.dtype.kind in "fi" and x.dtype.itemsize in (4, 8):
        try:
            import torch

            if torch.cuda.is_available():
                xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
                if x.dtype.kind != "f" or bool(torch.isfinite(xd).all()):  # NaN ordering: leave it to NumPy
                    out = torch.unique(xd, sorted=True, return_inverse=return_inverse, return_counts=return_counts)
                    out = out if isinstance(out, tuple) else (out,)
                    res = tuple(t.cpu().numpy() for t in out)
                    res = (res[0].astype(x.dtype, copy=False),) + tuple(r.astype(np.intp, copy=False) for r in res[1:])
                    return res if len(res) > 1 else res[0]
        except ImportError:  # pragma: no cover
            pass
    return np.unique(x, return_inverse=return_inverse, return_counts=return_counts)


def hist_quantized_ecdf(
    x: np.ndarray,
    *,
    density: bool = False,
    max_bin_error: float = 0.0125,
    max_bin_size: float = 0.125,
    merge_bin_size: float = 0.025,
    _values_counts=None,
    _n=None,
):
    """Variable-width histogram of `x` obtained by quantising its ECDF.  Returns (hist, bin_edges)."""
    n = len(x) if _n is None else _n
    err_abs, size_abs, merge_abs = int(max_bin_error * n), int(max_bin_size * n), int(merge_bin_size * n)
    values, counts = _values_counts if _values_counts is not None else unique_values(x, return_counts=True)
    cum = np.cumsum(counts)
    xs = np.insert(np.append(values, np.inf), 0, -np.inf)
    ys = np.insert(np.append(cum, np.iinfo(cum.dtype).max), 0, 0)
    lo, hi = 1, len(xs) - 1
    edges_lo, edges_hi = [values[0]], [values[-1]]
    hist_lo: list = []
    hist_hi: list = []
    hist, edges = [], []
    while lo < hi:
        lo_before, hi_before = lo, hi
        lo, n_lo = _grow_bin(xs, ys, lo, err_abs, size_abs, True)
        hi, n_hi = _grow_bin(xs, ys, hi, err_abs, size_abs, False)
        hist_lo.append(n_lo)
        hist_hi.insert(0, n_hi)
        edges_lo.append((xs[lo] + xs[lo - 1]) / 2 if lo > 0 else xs[lo])
        edges_hi.insert(0, (xs[hi] + xs[hi - 1]) / 2 if hi > 0 else xs[hi])
        if lo == hi:  # the two fronts met exactly
            edges = edges_lo + edges_hi[1:]
            hist = hist_lo + hist_hi
            break
        if lo > hi:  # the fronts crossed: the last two bins overlap, fuse them
            middle = cum[-1] - np.sum(hist_lo[:-1]) - np.sum(hist_hi[1:])
            hist = hist_lo[:-1] + [middle] + hist_hi[1:]
            edges = edges_lo[:-1] + edges_hi[1:]
            break
        if ys[hi - 1] - ys[lo - 1] <= merge_abs:  # small remainder: split it between the neighbours
            mid_lo = int(np.floor((lo + hi) / 2))
            mid_hi = int(np.ceil((lo + hi) / 2))
            centre = (xs[mid_lo] + xs[mid_hi]) / 2
            hist = (
                hist_lo[:-1]
                + [ys[mid_lo] - ys[lo_before - 1]]
                + [ys[hi_before - 1] - ys[mid_hi - 1]]
                + hist_hi[1:]
            )
            edges = edges_lo[:-1] + [centre] + edges_hi[1:]
            break
    fdtype = values.dtype if np.issubdtype(values.dtype, np.floating) else np.float64
    hist_arr = (np.array(hist) / cum[-1]).astype(fdtype) if density else np.array(hist)
    return hist_arr, np.array(edges).astype(fdtype)


class Quantizer(BaseEstimator, TransformerMixin):
    """Maps each numeric column to the index of its ECDF-quantised bin (reference :180-243)."""

    def __init__(self, *, max_bin_error: float = 0.0125, max_bin_size: float = 0.125,
                 append_invfreq: bool = False, dtype: Any = np.intp):
        self.max_bin_error = max_bin_error
        self.max_bin_size = max_bin_size
        self.append_invfreq = append_invfreq
        self.dtype = dtype
        if append_invfreq and not np.issubdtype(dtype, np.floating):
            self.dtype = np.float32

    def fit(self, X, y=None):
        X = check_array(X)
        self.n_features_in_ = X.shape[1]
        self.X_hist_, self.X_bin_edges_ = [], []
        for j in range(X.shape[1]):
            hist, edges = hist_quantized_ecdf(
                X[:, j], density=False, max_bin_error=self.max_bin_error, max_bin_size=self.max_bin_size
            )
            self.X_hist_.append(hist)
            self.X_bin_edges_.append(edges)
        return self

    def transform(self, X):
        ncol = X.shape[1]
        out = np.empty((X.shape[0], (1 + self.append_invfreq) * ncol), dtype=self.dtype)
        for j in range(ncol):
            edges = self.X_bin_edges_[j]
            code = np.clip(np.searchsorted(edges, X[:, j], side="right") - 1, 0, len(edges) - 2)
            out[:, j] = code
            if self.append_invfreq:
                out[:, ncol + j] = 1 / len(self.X_hist_[j]) / self.X_hist_[j][code]
        return out


def _device_sample_bins(x: np.ndarray, **kwargs: Any):
    """`sample_bins_quantized_ecdf` for a large vector with the sort, the counts AND the final binning on the device: the
    rank codes never come back to the host (their `searchsorted` against the bin edges alone cost 0.3 s at n = 4M), only
    the k counts for the ECDF scan and the n bin indices do.  Returns None when the device route does not apply."""
    x = np.asarray(x)
    if not (x.ndim == 1 and x.size >= MIN_SAMPLES_FOR_DEVICE_UNIQUE and x.dtype.kind in "fi" and x.dtype.itemsize in (4, 8)):
        return None
    if os.environ.get("NLS_HOST_BINS") == "1":  # A/B switch for timing the host recipe
        return None
    try:
        import torch
    except ImportError:  # pragma: no cover
        return None
    if not torch.cuda.is_available():
        return None
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    if x.dtype.kind == "f" and not bool(torch.isfinite(xd).all()):  # NaN ordering: leave it to NumPy
        return None
    distinct, codes, counts = torch.unique(xd, sorted=True, return_inverse=True, return_counts=True)
    k = int(distinct.numel())
    if k <= np.ceil(np.sqrt(x.size)):
        return codes.cpu().numpy().astype(np.intp, copy=False)
    q = Quantizer(dtype=np.intp, **kwargs)
    _, edges = hist_quantized_ecdf(
        None, density=False, max_bin_error=q.max_bin_error, max_bin_size=q.max_bin_size, _n=x.size,
        _values_counts=(np.arange(k, dtype=np.intp), counts.cpu().numpy().astype(np.intp, copy=False)),
    )
    # Quantizer.transform on the device: searchsorted(edges, code, side="right") - 1, clipped to the bins
    edges_d = torch.from_numpy(np.ascontiguousarray(edges, dtype=np.float64)).cuda()
    bins = torch.bucketize(codes.to(torch.float64), edges_d, right=True) - 1
    return bins.clamp_(0, len(edges) - 2).cpu().numpy().astype(np.intp, copy=False)


def sample_bins_quantized_ecdf(x: np.ndarray, **kwargs: Any) -> np.ndarray:
    """Bin index per sample: the class code if there are few distinct values, else ECDF bins (:246-253)."""
    on_device = _device_sample_bins(x, **kwargs)
    if on_device is not None:
        return on_device
    distinct, codes = unique_values(x, return_inverse=True)
    if len(distinct) <= np.ceil(np.sqrt(len(codes))):
        return codes
    # The codes are dense ranks 0..k-1: their distinct values and counts are a bincount, not another sort.
    q = Quantizer(dtype=np.intp, **kwargs)
    q.n_features_in_ = 1
    hist, edges = hist_quantized_ecdf(
        codes, density=False, max_bin_error=q.max_bin_error, max_bin_size=q.max_bin_size,
        _values_counts=(np.arange(len(distinct), dtype=codes.dtype), np.bincount(codes, minlength=len(distinct))),
    )
    q.X_hist_, q.X_bin_edges_ = [hist], [edges]
    return q.transform(codes[:, np.newaxis]).ravel()


def sample_weights_quantized_ecdf(x: np.ndarray, **kwargs: Any) -> np.ndarray:
    """Inverse-frequency sample weights from the same quantisation (:256-264)."""
    fdtype = x.dtype if np.issubdtype(x.dtype, np.floating) else np.float64
    distinct, codes, counts = np.unique(x, return_inverse=True, return_counts=True)
    if len(distinct) <= np.ceil(np.sqrt(len(codes))):
        return counts[codes] / np.sum(counts)
    return Quantizer(append_invfreq=True, dtype=fdtype, **kwargs).fit_transform(codes[:, np.newaxis])[:, 1]
