"""`NeoLSSVM`: the sklearn-compatible estimator, with the numerical core on a B200.

Drop-in mirror of the reference estimator (/root/reference/src/neo_ls_svm/_neo_ls_svm.py:43-821): same
keyword-only constructor, same public methods (`fit`, `decision_function`, `predict`, `predict_proba`,
`predict_std`, `predict_quantiles`, `predict_interval`, `score`) and the same fitted attributes
(`β̂_ γ_ γs_ loo_errors_γs_ loo_residuals_ loo_ŷ_ loo_leverage_ loo_error_ loo_score_ L_ residuals_
loo_std_ primal_feature_map_ dual_feature_map_ X_ α̂_ dual_ primal_ classes_ y_dtype_ n_features_in_
predict_proba_calibrator_ *_calib_l1_/_l2_ conformal_l1_/_l2_`), all host NumPy so that `pickle` and
`sklearn.base.clone` behave as before.

What moved to the GPU: the feature map, the Hermitian Gram, the eigendecomposition,
the leave-one-out γ sweep, the per-row LOO outputs, `decision_function`, `predict_std` and the
per-row part of `predict_quantiles` (primal), and the kernel-matrix / eigen / LOO pipeline of the dual
solve.  What stays on the host: input validation, the supervised affine pre-pass
(`AffineSeparator.fit`), isotonic calibration, the conformal split and the two small quantile LPs.

Host↔device traffic per primal fit: X, y, s up (once); A (m×m) and five n-vectors down.
"""

from __future__ import annotations

import os

from typing import Any, Literal

import numpy as np
import sklearn
from sklearn.base import BaseEstimator, clone
from sklearn.isotonic import IsotonicRegression
from sklearn.metrics import accuracy_score, r2_score
from sklearn.model_selection import train_test_split
from sklearn.utils import assert_all_finite
from sklearn.utils.validation import check_array, check_consistent_length, check_is_fitted, check_X_y

from . import _affine
from ._affine import AffineSeparator
from ._clqr import CoherentLinearQuantileRegressor
from ._feature_maps import KernelApproximatingFeatureMap, OrthogonalRandomFourierFeatures
from ._quantizer import unique_values

_DEVICE_FINITE_SCAN_MIN = 1 << 24  # elements of X above which the NaN/inf scan runs on the device copy

_DEVICE_STATE = "_device_state"
_SHARDED_PREDICT_MIN_ROWS = 1 << 16  # below this one GPU answers faster than the shards can be dealt out


def _is_frame(obj) -> bool:
    return hasattr(obj, "dtypes") and hasattr(obj, "index")


def _clip_correct_side(res: np.ndarray, y: np.ndarray) -> None:
    res[(y > 0) & (res > 0)] = 0
    res[(y < 0) & (res < 0)] = 0


_THREADED_SPLIT_MIN_ROWS = 1 << 18
_ASYNC_UPLOAD_MIN_BYTES = 64 << 20


def _shuffle_split(*arrays, train_size: int, random_state):
    """`sklearn.model_selection.train_test_split(*arrays, train_size=train_size, random_state=random_state)` for 1-D
    NumPy vectors: the same permutation from the same generator (ShuffleSplit._iter_indices: test rows first, then the
    train rows) and therefore the same outputs, with the eight gathers spread over host threads — at n = 4M the
    library call spends 0.18 s of the 2.7 s fit walking four 32 MB vectors through a random permutation one after the
    other.  Small inputs go to the library itself."""
    n = len(arrays[0])
    if n < _THREADED_SPLIT_MIN_ROWS or any(not isinstance(a, np.ndarray) or a.ndim != 1 or len(a) != n for a in arrays):
        return train_test_split(*arrays, train_size=train_size, random_state=random_state)
    from concurrent.futures import ThreadPoolExecutor

    from sklearn.utils import check_random_state

    n_train = int(train_size)
    n_test = n - n_train
    assert 0 < n_train < n
    perm = check_random_state(random_state).permutation(n)
    test, train = perm[:n_test], perm[n_test : n_test + n_train]
    # np.take releases the GIL; the big (test) gathers are cut into pieces so all host threads help
    pieces = max(1, min(8, (os.cpu_count() or 1) // len(arrays)))
    bounds = np.linspace(0, n_test, pieces + 1).astype(np.intp)
    out_test = [np.empty(n_test, dtype=a.dtype) for a in arrays]

    def gather(k, i):
        np.take(arrays[k], test[bounds[i] : bounds[i + 1]], out=out_test[k][bounds[i] : bounds[i + 1]])

    with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, pieces * len(arrays)))) as pool:
        list(pool.map(lambda ki: gather(*ki), [(k, i) for k in range(len(arrays)) for i in range(pieces)]))
    out = []
    for k, a in enumerate(arrays):
        out.extend((a[train], out_test[k]))
    return out


class NeoLSSVM(BaseEstimator):
    """Neo LS-SVM (primal random-feature or dual kernel least-squares SVM with closed-form LOO tuning)."""

    def __init__(
        self,
        *,
        primal_feature_map: KernelApproximatingFeatureMap | Literal["auto"] = "auto",
        dual_feature_map: AffineSeparator | Literal["auto"] = "auto",
        dual: bool | Literal["auto"] = "auto",
        estimator_type: Literal["auto", "classifier", "regressor"] = "auto",
        random_state: int | np.random.RandomState | None = 42,
    ) -> None:
        self.primal_feature_map = primal_feature_map
        self.dual_feature_map = dual_feature_map
        self.dual = dual
        self.random_state = random_state
        self.estimator_type = estimator_type

    # ------------------------------------------------------------------------------------------
    # pickling: device buffers are a cache, never part of the fitted state
    # ------------------------------------------------------------------------------------------
    def __getstate__(self):
        state = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        state = dict(state)
        state.pop(_DEVICE_STATE, None)
        return state

    # ------------------------------------------------------------------------------------------
    # device helpers
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _gpu():
        import torch

        from . import _lib, _multi

        ctx = _lib.context(_multi.devices()[0])  # the first selected GPU hosts the replicated solves and the pre-pass
        return ctx, torch, torch.device("cuda", ctx.device)

    def _primal_device_state(self, want_std: bool = False):
        """Device copies of what predict needs (rebuilt lazily, e.g. after unpickling)."""
        st = self.__dict__.get(_DEVICE_STATE)
        ctx, torch, dev = self._gpu()
        if st is None or st.get("kind") != "primal":
            shift, W = self.primal_feature_map_.device_weights(self.n_features_in_)
            st = {
                "kind": "primal",
                "shift": torch.from_numpy(shift).to(dev),
                "W": torch.from_numpy(W).to(dev),
                "beta": torch.from_numpy(np.asarray(self.β̂_, dtype=np.complex128)).to(dev),
                "U": torch.from_numpy(np.asarray(self.L_[0], dtype=np.complex128)).to(dev),
            }
            self.__dict__[_DEVICE_STATE] = st
        if want_std and "B" not in st:
            # (γC + A)⁻¹ = U⁻¹ U⁻ᴴ: the variance kernel takes the triangular B = U⁻¹ with unit weights and
            # skips the zero half of the contraction (4m² instead of 8m² flops per row).  Built once, lazily.
            m = st["U"].shape[0]
            st["B"] = ctx.triangular_inverse(st["U"])  # hand-written blocked inverse (csrc/potrf.cuh), upper triangle only
            st["w"] = torch.ones(m, dtype=torch.float64, device=dev)
        return st

    # ------------------------------------------------------------------------------------------
    # solvers
    # ------------------------------------------------------------------------------------------
    def _optimize_β̂_γ(self, X, y, s):
        """GPU counterpart of the reference's `_optimize_β̂_γ` (:77-189); takes X, not the materialised φ."""
        from . import _primal

        ctx, torch, dev = self._gpu()
        self._check_primal_feature_map(self.primal_feature_map_)
        shift, W = self.primal_feature_map_.device_weights(X.shape[1])
        dt = X.dtype
        s64 = np.asarray(s, dtype=np.float64)
        s_norm = s64 / np.sum(s64)  # :110
        Xd = _affine.device_copy(X)  # uploaded once in `fit`, shared with the supervised pre-pass
        if Xd is None:
            Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
        yd = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
        sd = torch.from_numpy(s_norm).to(dev)
        shd, Wd = torch.from_numpy(shift).to(dev), torch.from_numpy(W).to(dev)
        from . import _multi

        devs = _multi.devices()
        if len(devs) > 1:  # rows sharded over the selected GPUs (NLS_DEVICES / set_devices), SURVEY.md §8e
            fit = _multi.primal_fit_sharded(X, y, s_norm, shift, W, self._estimator_type == "classifier", devs, X_primary=Xd)
        else:
            fit = _primal.primal_fit(Xd, yd, sd, shd, Wd, self._estimator_type == "classifier", ctx=ctx)
        del Xd
        cdt = np.complex64 if dt == np.float32 else np.complex128
        self.γs_ = fit.gammas.astype(dt)
        self.loo_errors_γs_ = fit.loo_errors.astype(dt)
        stacked = fit.rows["_stacked"].cpu().numpy()  # one device -> host copy for the five per-row vectors
        rows = dict(zip(("loo_residuals", "yhat_loo", "loo_leverage", "residuals", "loo_std"), stacked))
        self.loo_residuals_ = rows["loo_residuals"].astype(dt, copy=False)
        self.loo_ŷ_ = (np.asarray(y, dtype=np.float64) + rows["loo_residuals"]).astype(dt, copy=False)
        self.loo_leverage_ = rows["loo_leverage"].astype(dt, copy=False)
        self.loo_error_ = self.loo_errors_γs_[fit.opt]
        self.loo_score_ = fit.loo_score
        self.L_ = (fit.U.cpu().numpy().astype(cdt), False)  # scipy.linalg.cho_factor layout (:177)
        self.residuals_ = rows["residuals"].astype(dt, copy=False)
        self.loo_std_ = rows["loo_std"].astype(dt, copy=False)
        self.__dict__[_DEVICE_STATE] = {"kind": "primal", "shift": shd, "W": Wd, "beta": fit.beta, "U": fit.U}
        return fit.beta.cpu().numpy().astype(cdt), self.γs_[fit.opt]

    @staticmethod
    def _check_primal_feature_map(fm) -> None:
        """The device solver computes φ = [exp(-i (x - shift) W)/√D | 1] itself and regularises with C = c·I
        (the reference's `eigh(A / c)` branch, :119-121).  A user-supplied map with another `transform` or a
        non-constant / non-diagonal complexity matrix (the reference's generalised `eigh(A, C)` + LU branch,
        :122-124) would be solved as something it is not, so it is refused instead."""
        from ._feature_maps import RandomFourierFeatures

        if not isinstance(fm, RandomFourierFeatures) or type(fm).transform is not RandomFourierFeatures.transform:
            raise NotImplementedError(
                "the B200 primal solver supports RandomFourierFeatures / OrthogonalRandomFourierFeatures maps "
                f"(and subclasses that keep their transform); got {type(fm).__name__}")
        C = np.asarray(fm.complexity_matrix)
        c = np.diag(C) if C.ndim == 2 else C  # noqa: PLR2004
        if (C.ndim == 2 and np.any(C != np.diag(c))) or np.any(c != c[0]) or not np.isfinite(c[0]) or c[0] <= 0:  # noqa: PLR2004
            raise NotImplementedError(
                "the B200 primal solver supports a constant diagonal complexity matrix only (every feature map the "
                "reference ships has one, _feature_maps.py:129-135)")

    def _optimize_α̂_γ(self, X, y, s, ρ: float = 1.0):
        """GPU counterpart of the reference's `_optimize_α̂_γ` (:191-325) for ρ = 1."""
        from . import _dual

        assert ρ == 1.0, "only the default ρ = 1 of the reference is supported"
        return _dual.fit_into(self, X, y, s)

    # ------------------------------------------------------------------------------------------
    # fit
    # ------------------------------------------------------------------------------------------
    def fit(self, X, y, sample_weight=None) -> "NeoLSSVM":
        """Fit this predictor."""
        import time

        marks = [("start", time.perf_counter())]  # host wall-clock marks of the phases, kept in `fit_phases_`
        # A large X is scanned for NaN/inf on the device, after the upload it needs anyway (0.2 s on the host at
        # n = 4M, d = 64); everything else about the validation, and the error raised, is sklearn's.
        device_scan = isinstance(X, np.ndarray) and X.size >= _DEVICE_FINITE_SCAN_MIN and X.dtype in (np.float64, np.float32)
        with sklearn.config_context(assume_finite=device_scan or sklearn.get_config()["assume_finite"]):
            X, y = check_X_y(X, y, dtype=(np.float64, np.float32), ensure_min_samples=2)
        y = np.ravel(np.asarray(y))
        if device_scan and not sklearn.get_config()["assume_finite"]:
            if y.dtype.kind in "fc":
                assert_all_finite(y, input_name="y")
        else:
            device_scan = False
        self.n_features_in_ = X.shape[1]
        self.y_dtype_ = y.dtype
        sample_weight_ = (
            np.ones(y.shape, X.dtype) if sample_weight is None else np.ravel(np.asarray(sample_weight)).astype(X.dtype)
        )
        check_consistent_length(y, sample_weight_)
        # Task type from the target (:351-373).
        # For the inference only "exactly two distinct values" matters: three distinct values in a prefix settle it without
        # sorting all of a large target vector (0.1 s at n = 4M).  A forced classifier needs the full set for `classes_`.
        head = np.unique(y[:4096]) if len(y) > (1 << 20) and self.estimator_type != "classifier" else ()
        distinct = head if len(head) > 2 else unique_values(y)  # noqa: PLR2004
        inferred = None
        if len(distinct) == 2:  # noqa: PLR2004
            inferred = "classifier"
        elif any(np.issubdtype(y.dtype, t) for t in (np.number, np.datetime64, np.timedelta64)):
            inferred = "regressor"
        self._estimator_type = inferred if self.estimator_type == "auto" else self.estimator_type
        if self._estimator_type == "classifier":
            self.classes_ = distinct
            y_ = np.ones(y.shape, dtype=X.dtype)
            y_[y == self.classes_[0]] = -1
        elif self._estimator_type == "regressor":
            y_ = y.astype(X.dtype)
        else:
            raise ValueError("Target type not supported")
        marks.append(("validation", time.perf_counter()))
        self.dual_ = X.shape[0] <= 1024 if self.dual == "auto" else self.dual  # noqa: PLR2004
        self.primal_ = not self.dual_
        self.__dict__.pop(_DEVICE_STATE, None)
        if self.primal_:
            self.primal_feature_map_ = clone(
                OrthogonalRandomFourierFeatures() if self.primal_feature_map == "auto" else self.primal_feature_map
            )
            # One host→device copy of X serves the supervised affine pre-pass and the solver.
            ctx, torch, dev = self._gpu()
            marks.append(("feature_map_clone_and_context", time.perf_counter()))
            if X.nbytes >= _ASYNC_UPLOAD_MIN_BYTES:
                # the copy (and the finiteness scan behind it) runs on a side stream underneath the host part of the
                # supervised pre-pass; whoever asks for the device copy first waits for it and sees the scan's verdict
                pending = _affine.PendingUpload(X, dev, device_scan)
                _affine.register_device_copy(X, pending)
            else:
                Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
                if device_scan and not bool(torch.isfinite(Xd).all()):
                    del Xd
                    assert_all_finite(X, input_name="X")  # raises sklearn's ValueError
                _affine.register_device_copy(X, Xd)
                del Xd
            try:
                # X, y and the weights were validated above; the nested transformers re-validate the same
                # arrays, so their finiteness scans (0.7 s at n = 4M) are switched off for this scope.
                marks.append(("upload_started", time.perf_counter()))
                with sklearn.config_context(assume_finite=True):
                    self.primal_feature_map_.fit(X, y_, sample_weight_)
                marks.append(("feature_map_fit", time.perf_counter()))
                self.β̂_, self.γ_ = self._optimize_β̂_γ(X, y_, sample_weight_)
                marks.append(("solve", time.perf_counter()))
            finally:
                _affine.release_device_copy(X)
        else:
            if device_scan:
                assert_all_finite(X, input_name="X")
            keep = sample_weight_ > 0
            X, y_, sample_weight_ = X[keep], y_[keep], sample_weight_[keep]
            self.dual_feature_map_ = clone(AffineSeparator() if self.dual_feature_map == "auto" else self.dual_feature_map)
            self.dual_feature_map_.fit(X, y_, sample_weight_)
            self.X_ = self.dual_feature_map_.transform(X)
            self.α̂_, self.γ_ = self._optimize_α̂_γ(self.X_, y_, sample_weight_)
        # Isotonic probability calibration on the LOO predictions (:406-412).
        if self._estimator_type == "classifier":
            self.predict_proba_calibrator_ = IsotonicRegression(out_of_bounds="clip", y_min=0, y_max=1, increasing=True)
            target = np.zeros_like(y_)
            target[y_ == np.max(y_)] = 1.0
            self.predict_proba_calibrator_.fit(self.loo_ŷ_, target, sample_weight_)
        # Two-level conformal calibration split of the LOO predictions (:414-430).
        (
            self.nonconformity_calib_l1_, self.nonconformity_calib_l2_,
            self.ŷ_calib_l1_, self.ŷ_calib_l2_,
            self.residuals_calib_l1_, self.residuals_calib_l2_,
            self.sample_weight_calib_l1_, self.sample_weight_calib_l2_,
        ) = _shuffle_split(
            self.loo_std_, self.loo_ŷ_, self.loo_residuals_, sample_weight_,
            train_size=min(1440, max(1024, (X.shape[0] * 2) // 3), X.shape[0] - 1),
            random_state=self.random_state,
        )
        self.conformal_l1_ = {"Δŷ": {}, "Δŷ/ŷ": {}}
        self.conformal_l2_ = {"Δŷ": {}, "Δŷ/ŷ": {}}
        marks.append(("calibration_split", time.perf_counter()))
        self.fit_phases_ = {name: t - marks[i][1] for i, (name, t) in enumerate(marks[1:])}  # seconds per phase
        if self.primal_ and "pending" in locals():
            self.fit_phases_.update(upload_alloc=pending.alloc_s, upload_total=pending.total_s, upload_waited_for=pending.waited_s)
        return self

    # ------------------------------------------------------------------------------------------
    # point predictions and predictive standard deviation (one fused device pass over φ(x))
    # ------------------------------------------------------------------------------------------
    def _decision_and_std(self, X: np.ndarray, want_decision: bool, want_std: bool):
        dt = X.dtype
        ctx, torch, dev = self._gpu()
        if self.primal_:
            from . import _multi

            st = self._primal_device_state(want_std)
            devs = _multi.devices()
            if len(devs) > 1 and X.shape[0] >= _SHARDED_PREDICT_MIN_ROWS:  # independent rows: shard, no collective
                yhat, sigma = _multi.primal_predict_sharded(
                    X, st["shift"], st["W"], st["beta"], st.get("B"), st.get("w"), devs, want_decision, want_std)
                return (yhat.astype(dt) if yhat is not None else None), (sigma.astype(dt) if sigma is not None else None)
            Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
            yhat, sigma = ctx.primal_predict(
                Xd, st["shift"], st["W"], beta=st["beta"] if want_decision else None,
                B=st["B"] if want_std else None, w=st["w"] if want_std else None, want_std=want_std, b_upper=True,
            )
        else:
            from . import _dual

            yhat, sigma = _dual.predict(self, X, want_decision, want_std)
        yhat = yhat.cpu().numpy().astype(dt) if yhat is not None else None
        sigma = sigma.cpu().numpy().astype(dt) if sigma is not None else None
        return yhat, sigma

    def decision_function(self, X):
        """Evaluate the prediction function ŷ(x) (:655-681)."""
        check_is_fitted(self)
        X, X_df = check_array(X, dtype=(np.float64, np.float32)), X
        ŷ, _ = self._decision_and_std(X, True, False)
        if _is_frame(X_df):
            try:
                import pandas as pd
            except ImportError:
                pass
            else:
                return pd.Series(ŷ, index=X_df.index)
        return ŷ

    def predict_std(self, X):
        """Bayesian estimate of the predictive standard deviation (:452-487)."""
        check_is_fitted(self)
        X, X_df = check_array(X, dtype=(np.float64, np.float32)), X
        _, σ = self._decision_and_std(X, False, True)
        if _is_frame(X_df):
            try:
                import pandas as pd
            except ImportError:
                pass
            else:
                return pd.Series(σ, index=X_df.index)
        return σ

    # ------------------------------------------------------------------------------------------
    # conformal quantiles
    # ------------------------------------------------------------------------------------------
    def _lazily_fit_conformal_predictor(self, target_type: str, quantiles):
        """Level-1 coherent quantile regressor + level-2 conformal bias, cached per quantile tuple (:489-532)."""
        quantiles = np.asarray(quantiles)
        key = tuple(quantiles)
        if key in self.conformal_l1_[target_type]:
            return self.conformal_l1_[target_type][key], self.conformal_l2_[target_type][key]
        relative = "/ŷ" in target_type
        regressor = self._estimator_type == "regressor"
        eps = np.finfo(self.ŷ_calib_l1_.dtype).eps

        def design(nonconformity, ŷ):
            cols = [nonconformity[:, np.newaxis]]
            if regressor:
                cols.append(np.abs(ŷ[:, np.newaxis]))
            return np.hstack(cols) if len(cols) > 1 else cols[0]

        def target(residuals, ŷ):
            return -residuals / (np.maximum(np.abs(ŷ), eps) if relative else 1)

        X1, y1 = design(self.nonconformity_calib_l1_, self.ŷ_calib_l1_), target(self.residuals_calib_l1_, self.ŷ_calib_l1_)
        cqr = CoherentLinearQuantileRegressor(quantiles=quantiles)
        cqr.fit(X1, y1, sample_weight=self.sample_weight_calib_l1_)
        self.conformal_l1_[target_type][key] = cqr
        bias = np.zeros(quantiles.shape, dtype=self.ŷ_calib_l1_.dtype)
        if len(self.ŷ_calib_l2_) >= 128:  # noqa: PLR2004
            X2, y2 = design(self.nonconformity_calib_l2_, self.ŷ_calib_l2_), target(self.residuals_calib_l2_, self.ŷ_calib_l2_)
            pred2 = cqr.predict(X2)
            clip = cqr.intercept_clip(np.vstack([X1, X2]), np.hstack([y1, y2]))
            for j, q in enumerate(quantiles):
                bias[j] = np.clip(np.quantile(y2 - pred2[:, j], q), clip[0, j], clip[1, j])
        self.conformal_l2_[target_type][key] = bias
        return cqr, bias

    def predict_quantiles(self, X, *, quantiles=(0.025, 0.5, 0.975), priority: Literal["accuracy", "coverage"] = "accuracy"):
        """Predict conformally calibrated quantiles (:554-624)."""
        check_is_fitted(self)
        X, X_df = check_array(X, dtype=(np.float64, np.float32)), X
        ŷ, σ = self._decision_and_std(X, True, True)  # φ(x) is generated once for both
        cqr_abs, bias_abs = self._lazily_fit_conformal_predictor("Δŷ", quantiles)
        cqr_rel, bias_rel = self._lazily_fit_conformal_predictor("Δŷ/ŷ", quantiles)
        if priority == "coverage":  # only allow the quantiles to move outwards; mutates the cache as the reference does
            q = np.asarray(quantiles)
            for bias in (bias_abs, bias_rel):
                bias[0.5 <= q] = np.maximum(bias[0.5 <= q], 0)
                bias[q <= 0.5] = np.minimum(bias[q <= 0.5], 0)
        regressor = self._estimator_type == "regressor"
        ctx, torch, dev = self._gpu()

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

        iso = self.predict_proba_calibrator_ if not regressor else None
        out = ctx.quantile_epilogue(
            up(ŷ), up(σ), up(cqr_abs.β_), up(cqr_rel.β_), up(bias_abs), up(bias_rel), regressor,
            up(iso.X_thresholds_) if iso is not None else None, up(iso.y_thresholds_) if iso is not None else None,
        )
        ŷ_quantiles = out.cpu().numpy().astype(X.dtype)
        if regressor and not np.issubdtype(self.y_dtype_, np.integer):
            ŷ_quantiles = ŷ_quantiles.astype(self.y_dtype_)
        if _is_frame(X_df):
            try:
                import pandas as pd
            except ImportError:
                pass
            else:
                if regressor:
                    frame = pd.DataFrame(ŷ_quantiles, index=X_df.index, columns=quantiles)
                else:
                    neg = pd.DataFrame(ŷ_quantiles[:, :, 0], index=X_df.index, columns=quantiles)
                    pos = pd.DataFrame(ŷ_quantiles[:, :, 1], index=X_df.index, columns=quantiles)
                    frame = pd.concat([neg, pos], axis=0, keys=self.classes_, names=["class", X_df.index.name])
                frame.columns.name = "quantile"
                return frame
        return ŷ_quantiles

    def predict_interval(self, X, *, coverage: float = 0.95):
        """Predict conformally calibrated intervals (:636-645)."""
        lb = (1 - coverage) / 2
        return self.predict_quantiles(X, quantiles=(lb, 1 - lb), priority="coverage")

    # ------------------------------------------------------------------------------------------
    # predict / predict_proba / score
    # ------------------------------------------------------------------------------------------
    def predict(self, X, *, coverage: float | None = None, quantiles=None):
        """Predict on a given dataset (:719-762)."""
        assert coverage is None or quantiles is None
        if coverage is not None:
            return self.predict_interval(X, coverage=coverage)
        if quantiles is not None:
            return self.predict_quantiles(X, quantiles=quantiles)
        check_is_fitted(self)
        X, X_df = check_array(X, dtype=(np.float64, np.float32)), X
        ŷ = self.decision_function(X)
        if self._estimator_type == "classifier":
            side = np.sign(ŷ)
            side[side == 0] = -1  # ties go to the negative class
            ŷ = self.classes_[((side + 1) // 2).astype(np.intp)]
        if not np.issubdtype(self.y_dtype_, np.integer):
            ŷ = ŷ.astype(self.y_dtype_)
        if _is_frame(X_df):
            try:
                import pandas as pd
            except ImportError:
                pass
            else:
                return pd.Series(ŷ, index=X_df.index)
        return ŷ

    def predict_proba(self, X):
        """Isotonically calibrated class probabilities (classifier) or the point prediction (regressor) (:772-799)."""
        check_is_fitted(self)
        X, X_df = check_array(X, dtype=(np.float64, np.float32)), X
        ŷ = self.decision_function(X)
        if self._estimator_type == "classifier":
            p = self.predict_proba_calibrator_.transform(ŷ)
            proba = np.hstack([1 - p[:, np.newaxis], p[:, np.newaxis]])
        else:
            proba = ŷ if np.issubdtype(self.y_dtype_, np.integer) else ŷ.astype(self.y_dtype_)
        if _is_frame(X_df):
            try:
                import pandas as pd
            except ImportError:
                pass
            else:
                if self._estimator_type == "regressor":
                    return pd.Series(proba, index=X_df.index)
                return pd.DataFrame(proba, index=X_df.index, columns=self.classes_)
        return proba

    def score(self, X, y, sample_weight=None) -> float:
        """Accuracy (classifier) or R² (regressor) (:801-817)."""
        ŷ = self.predict(X)
        if self._estimator_type == "classifier":
            return accuracy_score(y, ŷ, sample_weight=sample_weight)
        return r2_score(np.asarray(y).astype(np.float64), np.asarray(ŷ).astype(np.float64), sample_weight=sample_weight)

    def _more_tags(self) -> dict[str, Any]:
        return {"binary_only": True, "requires_y": True}

    def __sklearn_tags__(self):
        tags = super().__sklearn_tags__()
        tags.target_tags.required = True
        # Like the reference, the task type is only known once `fit` has seen the target (`_estimator_type`
        # is set there, :361-363); an unfitted instance is a generic estimator for sklearn's checks.
        est = self.__dict__.get("_estimator_type")
        if est in ("classifier", "regressor"):
            tags.estimator_type = est
        if est == "classifier":
            from sklearn.utils import ClassifierTags

            tags.classifier_tags = ClassifierTags(multi_class=False)
        elif est == "regressor":
            from sklearn.utils import RegressorTags

            tags.regressor_tags = RegressorTags()
        return tags
