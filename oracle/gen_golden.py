"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    NUMBA_CACHE_DIR=/tmp/nb PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [case ...]

The reference (/root/reference, read-only) cannot travel to the GPU box, so its outputs on the
seeded synthetic cases of `neo_ls_svm_b200.datasets.CASES` are committed as fixtures.  Each fixture
stores the fitted affine/Fourier map (so hot-path parity can be tested independently of the host
pre-pass), every fitted attribute of the solver, and predict/predict_std/predict_quantiles outputs
on held-out rows.  TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nb_cache")
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

from neo_ls_svm_b200.datasets import CASES, golden_row_subset, load_case  # noqa: E402


def run_case(name: str) -> dict:
    from neo_ls_svm import NeoLSSVM
    from neo_ls_svm._feature_maps import OrthogonalRandomFourierFeatures

    X, y, sw, Xt, est = load_case(name)
    kwargs = {}
    if "num_features" in est:
        kwargs["primal_feature_map"] = OrthogonalRandomFourierFeatures(num_features=est["num_features"])
    if "dual" in est:
        kwargs["dual"] = est["dual"]
    model = NeoLSSVM(**kwargs).fit(X, y, sample_weight=sw)
    rows = golden_row_subset(len(y))  # per-row vectors of large cases are stored on a fixed row subset
    out: dict = {
        "rows": rows,
        "x_checksum": np.array([X.sum(), np.abs(X).sum(), float(np.asarray(y, dtype=np.float64).sum())]),
        "classifier": np.array(model._estimator_type == "classifier"),
        "dual": np.array(bool(model.dual_)),
        "gamma": np.array(model.γ_),
        "gammas": model.γs_,
        "loo_errors": model.loo_errors_γs_,
        "opt": np.array(int(np.argmin(np.abs(model.γs_ - model.γ_)))),
        "loo_residuals": model.loo_residuals_[rows],
        "loo_yhat": model.loo_ŷ_[rows],
        "loo_error": np.array(model.loo_error_),
        "loo_score": np.array(model.loo_score_),
        "residuals": model.residuals_[rows],
        "loo_std": model.loo_std_[rows],
    }
    # Digest of the Cholesky factor U (gamma*C + A = U^H U): its diagonal and U @ probe.
    U = np.triu(model.L_[0])
    probe = np.cos(np.arange(U.shape[0]) * 0.7) + 0.25
    out["L_diag"] = np.diag(U)
    out["L_probe"] = U @ probe
    if model.primal_:
        fm = model.primal_feature_map_
        aff = fm.affine_feature_map
        out.update(
            shift=aff.shift_, scale=aff.scale_, A_map=aff.A_,
            beta=model.β̂_, loo_leverage=model.loo_leverage_[rows],
        )
        # The feature map itself on a few rows (a1 + a2).
        out["phi_head"] = fm.transform(X[:16])
    else:
        aff = model.dual_feature_map_
        out.update(shift=aff.shift_, scale=aff.scale_, A_map=aff.A_, Xt_train=model.X_, alpha=model.α̂_)
        assert len(rows) == len(y), "dual fixtures keep every row"
    # Predictions on held-out rows.
    out["decision"] = model.decision_function(Xt)
    out["std"] = model.predict_std(Xt)
    out["predict"] = model.predict(Xt)
    quantiles = (0.025, 0.5, 0.975)
    out["quantiles_accuracy"] = model.predict_quantiles(Xt, quantiles=quantiles)
    for kind, tag in (("Δŷ", "abs"), ("Δŷ/ŷ", "rel")):
        cqr = model.conformal_l1_[kind][tuple(np.asarray(quantiles))]
        out[f"cqr_{tag}_beta"] = cqr.β_
        out[f"cqr_{tag}_bias"] = model.conformal_l2_[kind][tuple(np.asarray(quantiles))].copy()
    out["interval_90"] = model.predict_interval(Xt, coverage=0.9)
    if model._estimator_type == "classifier":
        out["proba"] = model.predict_proba(Xt)
        cal = model.predict_proba_calibrator_
        out["iso_x"] = cal.X_thresholds_
        out["iso_y"] = cal.y_thresholds_
        out["classes"] = model.classes_
    # Calibration split (host post-step; pins train_test_split parity).
    out["calib_l1_head"] = model.ŷ_calib_l1_[:16]
    out["calib_l2_head"] = model.ŷ_calib_l2_[:16]
    return {k: np.asarray(v) for k, v in out.items()}


def bench_map() -> dict:
    """(shift, W) of the C3 benchmark: the reference's OrthogonalRandomFourierFeatures(1024) fitted on the first
    100,000 rows of the bench dataset, so that bench.py's two arms share one map and the reference arm never has to
    import the product package."""
    from neo_ls_svm._feature_maps import OrthogonalRandomFourierFeatures

    from neo_ls_svm_b200.datasets import fast_regression_rows

    X, y = fast_regression_rows(4_000_000, 64, 32, row_begin=0, row_end=100_000)
    fm = OrthogonalRandomFourierFeatures(num_features=1024).fit(X, y, np.ones(len(y)))
    aff = fm.affine_feature_map
    W = aff.A_ / np.reshape(aff.scale_, (-1, 1))
    return {"shift": np.ravel(aff.shift_), "W": np.ascontiguousarray(W), "rows": np.array(100_000),
            "x_checksum": np.array([X.sum(), np.abs(X).sum(), y.sum()])}


def main() -> None:
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    if "bench_map" in names:
        names.remove("bench_map")
        path = os.path.join(ROOT, "tests", "golden", "bench_c3_map.npz")
        np.savez_compressed(path, **bench_map())
        print("bench_map ->", path, f"{os.path.getsize(path) / 1e6:.2f} MB")
    for name in names:
        arrays = run_case(name)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "gamma idx", int(arrays["opt"]))


if __name__ == "__main__":
    main()
