"""Affine feature maps: `AffineFeatureMap`, supervised `AffineNormalizer` and `AffineSeparator`.

Host-side (NumPy) mirrors of the reference's transformers — same constructor arguments, fitted
attributes (`shift_`, `scale_`, `A_`) and numerical recipe:

* AffineFeatureMap   /root/reference/src/neo_ls_svm/_affine_feature_map.py:17-136
* AffineNormalizer   /root/reference/src/neo_ls_svm/_affine_normalizer.py:16-117
* AffineSeparator    /root/reference/src/neo_ls_svm/_affine_separator.py:54-210

Fitting them is the pre-pass *before* the five GPU stages and is out of the hot path's
scope for this round (§8f "next" #1): it stays on the host so that `shift_/scale_/A_` — and therefore
the matrix W handed to the kernels — match the reference for identical seeds.  `transform` of the
fitted map is stage 1 of the hot path and runs on the GPU inside the estimator; the NumPy `transform`
here exists for API parity on small inputs (e.g. the dual path's n ≤ 1024 rows).
"""

from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from functools import cached_property
from typing import Any

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array, check_consistent_length, check_random_state, check_X_y
from sklearn.utils.validation import _check_feature_names_in

from ._quantizer import sample_bins_quantized_ecdf
from ._weighted_quantile import weighted_quantile


class AffineFeatureMap(BaseEstimator, TransformerMixin):
    """x -> (x - shift) diag(1/scale) A   (A optional)."""

    def __init__(self, *, scale, shift, A=None, append_features: bool = False):
        self.scale = scale
        self.shift = shift
        self.A = A
        self.append_features = append_features

    # -- helpers -------------------------------------------------------------------------------
    def _params(self, n_features: int):
        scale = np.reshape(getattr(self, "scale_", self.scale), (-1, n_features))
        shift = np.reshape(getattr(self, "shift_", self.shift), (-1, n_features))
        return scale, shift, getattr(self, "A_", self.A)

    def device_weights(self, n_features: int, dtype=np.float64):
        """(shift, W) with z = (x - shift) W, i.e. W = A/scaleᵀ (or diag(1/scale) if A is None)."""
        scale, shift, A = self._params(n_features)
        W = (np.eye(n_features, dtype=dtype) if A is None else A) / scale.T
        return np.ascontiguousarray(shift.ravel(), dtype=dtype), np.ascontiguousarray(W, dtype=dtype)

    # -- sklearn API ---------------------------------------------------------------------------
    def fit(self, X, y=None, sample_weight=None):
        X = check_array(X)
        self.n_features_in_ = X.shape[1]
        scale, shift, A = self._params(X.shape[1])
        assert scale.dtype == shift.dtype, "The scale and shift must have the same dtype"
        assert not np.any(scale == 0), "The scale may not be zero"
        assert np.all(np.isfinite(scale)), "The scale must be finite"
        assert np.all(np.isfinite(shift)), "The shift must be finite"
        assert X.shape[1] == scale.shape[1], "The scale must be compatible with the number of features"
        assert X.shape[1] == shift.shape[1], "The shift must be compatible with the number of features"
        if A is not None:
            assert A.dtype == scale.dtype, "The matrix A must have the same dtype as the scale and shift"
            assert X.shape[1] == A.shape[0], "The matrix A must have rows equal to the number of features in X"
            assert np.all(np.isfinite(A)), "The matrix A must be finite"
        return self

    def transform(self, X):
        X = check_array(X)
        scale, shift, A = self._params(X.shape[1])
        if A is None:
            Z = (X - shift) / scale
        else:
            W = A / scale.T
            # Same association as the reference (:85-87): project first when A narrows the data.
            Z = X @ W - shift @ W if A.shape[1] < A.shape[0] else (X - shift) @ W
        Z = Z.astype(X.dtype)
        if self.append_features and A is not None:
            Z = np.hstack((X, Z))
        return Z

    @cached_property
    def pseudo_inverse(self):
        return np.linalg.pinv(self.A) if self.A is not None else None

    def inverse_transform(self, X_transformed):
        X = check_array(X_transformed)
        A = getattr(self, "A_", self.A)
        scale = np.reshape(getattr(self, "scale_", self.scale), (-1, X.shape[1] if A is None else A.shape[0]))
        shift = np.reshape(getattr(self, "shift_", self.shift), scale.shape)
        if self.append_features and A is not None:
            return X[:, : A.shape[0]]
        if A is not None:
            X = X @ self.pseudo_inverse
        return (X * scale + shift).astype(X.dtype)

    def get_feature_names_out(self, input_features=None):
        A = getattr(self, "A_", self.A)
        names = np.asarray(_check_feature_names_in(self, input_features), dtype=object)
        if A is None:
            out = names + "_shifted_scaled"
        else:
            out = np.array([f"{','.join(list(names))}_affine_map"] * A.shape[1], dtype=object)
        if self.append_features and A is not None:
            out = np.hstack((names, out))
        return out

    def _more_tags(self) -> dict[str, Any]:
        return {"preserves_dtype": [np.float64, np.float32]}


def _target_bins(y, sample_weight):
    """Group rows by the quantised target.

    Returns (rows per bin, total weight per bin, normalised row weights per bin as 1×n_b arrays).  Only
    index and weight vectors are built here; feature rows are gathered per bin when needed, so no second
    copy of X is ever alive.
    """
    codes = sample_bins_quantized_ecdf(y)
    rows = [np.flatnonzero(codes == c) for c in range(np.min(codes), np.max(codes) + 1)]
    mass = [np.sum(sample_weight[r]) for r in rows]
    s_bins = [sample_weight[np.newaxis, r] / np.sum(sample_weight[r]) for r in rows]
    return rows, mass, s_bins


# Device copies of training matrices registered by the estimator (key: host data pointer + shape), so the
# pre-pass and the solver share one upload of X.
_DEVICE_COPIES: dict = {}


def _array_key(X) -> tuple:
    return (X.__array_interface__["data"][0], X.shape, X.dtype.str)


class PendingUpload:
    """Host→device copy of a large training matrix on a side stream, driven by a host thread, so that the 0.2 s a
    pageable 2 GB upload takes runs underneath the host part of the supervised pre-pass (target quantisation) instead
    of in front of it.  `result()` joins, orders the caller's stream after the copy and reports non-finite input."""

    def __init__(self, X, device, scan_finite: bool):
        import threading

        import torch

        self._torch, self._X, self._scan, self._device = torch, X, scan_finite, device
        self._caller_stream = torch.cuda.current_stream(device)
        self._Xd = self._finite = self._event = self._error = self._stream = None
        self.alloc_s = self.total_s = self.waited_s = 0.0
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self) -> None:
        torch = self._torch
        try:
            import time

            t0 = time.perf_counter()
            self._stream = torch.cuda.Stream(self._device)
            self._stream.wait_stream(self._caller_stream)  # the block may be recycled from work queued by the caller
            with torch.cuda.stream(self._stream):
                # allocated here, on the side stream: a fresh 2 GB cudaMalloc is not free either
                self._Xd = torch.empty(self._X.shape, dtype=torch.float64, device=self._device)
                self.alloc_s = time.perf_counter() - t0
                src = torch.from_numpy(np.ascontiguousarray(self._X))
                # In pieces of ~64 MB: the copy engine serves the streams in issue order, and the pre-pass running on
                # the caller's stream has small transfers of its own (the target vector) that must not queue behind 2 GB.
                step = max(1, (64 << 20) // max(1, src.shape[1] * src.element_size()))
                for r0 in range(0, src.shape[0], step):
                    piece = src[r0 : r0 + step]
                    if src.dtype == torch.float64:
                        self._Xd[r0 : r0 + step].copy_(piece)
                    else:  # float32 rows cross the bus as they are and are widened on the device
                        self._Xd[r0 : r0 + step].copy_(piece.to(self._device))
                if self._scan:
                    self._finite = torch.isfinite(self._Xd).all()
                self._event = self._stream.record_event()
                self._event.synchronize()
                self.total_s = time.perf_counter() - t0
        except BaseException as exc:  # noqa: BLE001  (re-raised by result() on the caller's thread)
            self._error = exc

    def result(self):
        if self._thread is not None:
            import time

            t0 = time.perf_counter()
            self._thread.join()
            self.waited_s = time.perf_counter() - t0  # how long the consumer stood still for the copy
            self._thread = None
            if self._error is not None:
                raise self._error
            self._torch.cuda.current_stream(self._device).wait_event(self._event)
            self._Xd.record_stream(self._torch.cuda.current_stream(self._device))  # allocated on the side stream
            if self._scan and not bool(self._finite):
                import sklearn
                from sklearn.utils import assert_all_finite

                with sklearn.config_context(assume_finite=False):  # the caller may have switched re-validation off
                    assert_all_finite(self._X, input_name="X")  # raises sklearn's ValueError
        return self._Xd

    def abandon(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None


def register_device_copy(X, Xd) -> None:
    """Xd: the device tensor, or a PendingUpload that will produce it."""
    _DEVICE_COPIES[_array_key(X)] = Xd


def release_device_copy(X) -> None:
    held = _DEVICE_COPIES.pop(_array_key(X), None)
    if isinstance(held, PendingUpload):
        held.abandon()


def device_copy(X):
    held = _DEVICE_COPIES.get(_array_key(X))
    return held.result() if isinstance(held, PendingUpload) else held


def _bin_location_spread(X, rows, s_bins):
    """Per-bin weighted median and weighted mean absolute deviation of every feature.

    Large inputs go through the sort-free GPU kernels (`_binstats.py`, `nls_bin_median_stats`); small
    ones, and machines without a GPU, use the reference's host recipe (argsort per column per bin).
    """
    from . import _binstats

    Xd = device_copy(X)
    if Xd is None and X.size >= _binstats.MIN_ELEMENTS_FOR_DEVICE:
        try:
            import torch

            if torch.cuda.is_available():
                Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).cuda()
        except ImportError:  # pragma: no cover
            Xd = None
    if Xd is not None and X.size >= _binstats.MIN_ELEMENTS_FOR_DEVICE:
        centre, spread = _binstats.device_bin_location_spread(Xd, rows, s_bins)
        return [c.astype(X.dtype) for c in centre], [sp.astype(X.dtype) for sp in spread]
    centre, spread = [], []
    for r, sb in zip(rows, s_bins):
        Xb = X[r, :]
        mu = weighted_quantile(Xb, sb.T, 0.5, axis=0)
        centre.append(mu)
        spread.append(sb @ np.abs(Xb - mu))
    return centre, spread


class AffineNormalizer(AffineFeatureMap):
    """Supervised shift/scale: centre between, and scale by the spread of, the target's class bins."""

    def __init__(self, *, append_features: bool = False) -> None:
        self.shift = 0.0
        self.scale = 1.0
        self.A = None
        self.append_features = append_features

    def fit(self, X, y=None, sample_weight=None, _bins=None):
        X, y = check_X_y(X, y, dtype=(np.float64, np.float32))
        y = np.ravel(np.asarray(y)).astype(X.dtype)
        sw = (np.ones(y.shape) if sample_weight is None else np.ravel(np.asarray(sample_weight))).astype(y.dtype)
        check_consistent_length(y, sw)
        rows, mass, s_bins = _bins if _bins is not None else _target_bins(y, sw)
        d = X.shape[1]
        if len(rows) <= 1:
            self.shift_ = np.zeros((1, d), dtype=X.dtype)
            self.scale_ = np.ones((1, d), dtype=X.dtype)
            AffineFeatureMap.fit(self, X, y, sw)
            return self
        # Per-bin robust location (weighted median) and spread (weighted mean absolute deviation).
        centre, spread = _bin_location_spread(X, rows, s_bins)
        eps = np.finfo(X.dtype).eps
        direction = np.zeros((1, d), dtype=X.dtype)
        weight_sum = np.zeros((1, d), dtype=X.dtype)
        shift = np.zeros((1, d), dtype=X.dtype)
        scale = np.zeros((1, d), dtype=X.dtype)
        n_bins = len(centre)
        for i in range(n_bins - 1):
            for j in range(i + 1, n_bins):
                gap = centre[j] - centre[i]
                width = np.maximum(spread[i] + spread[j], eps)
                separability = np.abs(gap) / width
                # Pair weight: regularised geometric mean of the pair's mass and separability.
                w = np.sqrt((mass[i] + mass[j]) * (0.5 + separability))
                # Optimal threshold between the two bins, placed proportionally to their spreads.
                alpha = np.clip(spread[i] / width, 1e-6, 1.0 - 1e-6)
                shift = shift + w * (centre[i] + alpha * gap)
                scale = scale + w * width
                direction += w * np.sign(gap)
                weight_sum += w
        direction /= weight_sum
        self.shift_ = shift / weight_sum
        self.scale_ = scale / weight_sum
        flip = np.sign(direction) < 0
        self.scale_[flip] = -self.scale_[flip]
        AffineFeatureMap.fit(self, X, y, sw)
        return self


def _weight_cdf(p):
    """The normalised running sum `RandomState.choice` builds from p (cast to float64 first, as it does)."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf


def _cdf_lookup(cdf, uniforms):
    return np.searchsorted(cdf, uniforms, side="right")


def _weighted_draw(rng, p, size):
    """`rng.choice(len(p), size=size, p=p)` of a legacy RandomState, without its validation passes over p.

    Same algorithm (inverse-CDF lookup of `size` uniforms, numpy/random/mtrand.pyx) and therefore the same
    indices and the same generator state afterwards; p is a probability vector built a line earlier from
    validated sample weights, so re-checking it costs three more passes over up to n elements per call.
    """
    return _cdf_lookup(_weight_cdf(p), rng.random_sample(size))


def pairwise_distances(X, Y):
    """Squared Euclidean distances between the rows of X and Y."""
    return np.sum(X * X, axis=1, keepdims=True) - 2 * X @ Y.T + np.sum(Y * Y, axis=1, keepdims=True).T


def nearest_neighbours(X, Y):
    """For each row of X, the closest row of Y."""
    pick = np.argmin(pairwise_distances(X, Y), axis=1, keepdims=True)
    return np.take_along_axis(Y, pick, axis=0)


def _right_singular_vectors(X):
    """Singular values (descending) and right singular vectors of X via the smaller Gram matrix."""
    if X.shape[0] >= X.shape[1]:
        e, V = np.linalg.eigh(X.conj().T @ X)
        return np.sqrt(np.abs(e))[::-1], V[:, ::-1]
    e, U = np.linalg.eigh(X @ X.conj().T)
    sv = np.sqrt(np.abs(e))[::-1]
    U = U[:, ::-1]
    keep = sv > 0
    sv, U = sv[keep], U[:, keep]
    return sv, (X.conj().T @ U) / sv[np.newaxis, :]


class AffineSeparator(AffineNormalizer):
    """Supervised affine map: AffineNormalizer's shift/scale plus a matrix A that pulls the target's
    class bins apart, scaled for a unit-width Gaussian kernel (λ = sqrt(2 log(f/g)/(f-g)))."""

    def __init__(self, *, append_features: bool = False, rank_threshold: float = 2e-2,
                 edge_sample_size: int = 384, edge_search_multiplier: int = 4,
                 random_state: int | np.random.RandomState | None = 42) -> None:
        self.shift = 0.0
        self.scale = 1.0
        self.A = None
        self.append_features = append_features
        self.rank_threshold = rank_threshold
        self.edge_sample_size = edge_sample_size
        self.edge_search_multiplier = edge_search_multiplier
        self.random_state = random_state

    def fit(self, X, y=None, sample_weight=None):
        assert y is not None
        X, y = check_X_y(X, y, dtype=(np.float64, np.float32))
        y = np.ravel(np.asarray(y)).astype(X.dtype)
        sw = (np.ones(y.shape) if sample_weight is None else np.ravel(np.asarray(sample_weight))).astype(y.dtype)
        check_consistent_length(y, sw)
        bins = _target_bins(y, sw)  # the reference quantises y twice (normaliser and separator); once is enough
        AffineNormalizer.fit(self, X, y, sample_weight, _bins=bins)
        rows, mass, s_bins = bins
        n_bins = len(rows)
        if n_bins <= 1:
            return self
        if n_bins == 2:  # the reference enlarges (and keeps) the sample size for two bins (:139-141)
            self.edge_sample_size = int(self.edge_sample_size * 4 / 3)
        E, wide = self.edge_sample_size, self.edge_sample_size * self.edge_search_multiplier
        rng = check_random_state(self.random_state)
        shift = np.reshape(self.shift_, (1, -1))
        scale = np.reshape(self.scale_, (1, -1))

        def normalised(idx):
            # Rows of the shifted/scaled matrix (:121) without materialising it: the separator only ever
            # looks at a few thousand sampled rows.
            return ((X[idx, :] - shift) / scale).astype(X.dtype)

        sizes = np.array([len(r) for r in rows])
        sw_by_bin = [sw[r] for r in rows]  # gathered once; every bin's complement concatenates six of them
        # The three weighted draws per bin consume the generator in a fixed order (own seeds, complement, own sample)
        # and nothing else touches it in between, so the uniforms are drawn first, in that order, and the 2 n_bins
        # inverse-CDF lookups (a running sum over up to n weights each) then run side by side on the host threads.
        uniforms = [(rng.random_sample(E), rng.random_sample(wide), rng.random_sample(wide)) for _ in range(n_bins)]

        def draw_own(i):
            cdf = _weight_cdf(np.ravel(s_bins[i]))
            return _cdf_lookup(cdf, uniforms[i][0]), _cdf_lookup(cdf, uniforms[i][2])

        def draw_rest(i):
            w_rest = np.hstack([sw_by_bin[j] for j in range(n_bins) if j != i])
            return _cdf_lookup(_weight_cdf(np.ravel(w_rest) / np.sum(w_rest)), uniforms[i][1])

        with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, 2 * n_bins))) as pool:
            own_futures = [pool.submit(draw_own, i) for i in range(n_bins)]
            rest_futures = [pool.submit(draw_rest, i) for i in range(n_bins)]
            own_draws = [f.result() for f in own_futures]
            rest_draws = [f.result() for f in rest_futures]
        directions, inside_edges, outside_edges = [], [], []
        for i in range(n_bins):
            own_rows = rows[i]
            seeds = normalised(own_rows[own_draws[i][0]])
            # Sample the complement of bin i ("vstack of the other bins", :150-156) through row indices.
            others = [j for j in range(n_bins) if j != i]
            pick = rest_draws[i]
            offsets = np.concatenate([[0], np.cumsum(sizes[others])])
            which = np.searchsorted(offsets, pick, side="right") - 1
            rest_rows = np.array([rows[others[b]][k - offsets[b]] for b, k in zip(which, pick)])
            rest_sample = normalised(rest_rows)
            # Points of the complement closest to bin i, then points of bin i closest to those.
            outside = nearest_neighbours(seeds, rest_sample)
            own_sample = normalised(own_rows[own_draws[i][1]])
            inside = nearest_neighbours(outside, own_sample)
            outside_edges.append(outside)
            inside_edges.append(inside)
            sv, V = _right_singular_vectors(inside - outside)
            directions.append(V[:, : np.sum(sv > self.rank_threshold * sv[0])])
        self.A_ = np.hstack(directions)
        # Kernel-width scaling from the mean inter-bin (f) and intra-bin (g) squared edge distances.
        f = g = 0.0
        pairs_f, pairs_g = E * (E + 1) / 2, E * (E - 1) / 2
        for inside, outside, n_b in zip(inside_edges, outside_edges, mass, strict=True):
            pi, po = inside @ self.A_, outside @ self.A_
            f += n_b * np.sum(np.tril(pairwise_distances(pi, po), k=0)) / pairs_f
            g += n_b * np.sum(np.tril(pairwise_distances(pi, pi), k=-1)) / pairs_g
        f /= sum(mass)
        g /= sum(mass)
        lam = np.sqrt(2 * np.log(f / g) / (f - g)) if g > 0 else 1
        self.A_ *= lam
        return self
