"""Driver-visible parity for the BASELINE.json configurations that round 1 only covered in scripts:

* C2 at FULL size (n = 100,000, d = 70, classifier) against a fixture produced by the unmodified reference
  (`oracle/gen_golden.py c2_full`; per-row vectors on the 4096 rows of `datasets.golden_row_subset`), through the
  hot path with the reference's fitted map AND through the public `NeoLSSVM.fit` (device pre-pass, n·d ≥ 2²¹);
* a C5-shaped primal fit (d = 128, num_features = 4096 ⇒ m = 4097, `auto` eigensolver) against the chunked oracle;
* the dual path at n = 4096 (C4's code path beyond the n ≤ 1500 of the fixtures) against `orc.dual_fit`.

Tolerances are the north star's: same γ index; 1e-9 on β̂ / LOO residuals / error curve; 1e-7 on predictions.
`assert_elementwise` additionally bounds every entry: |Δ| ≤ 1e-9·|ref| + 1e-12·max|ref|.
"""

import numpy as np
import pytest

from conftest import assert_elementwise, rel_err  # noqa: E402
from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures
from neo_ls_svm_b200.datasets import golden_row_subset, load_case, make_regression_rows

pytestmark = pytest.mark.gpu

TOL_FIT, TOL_PRED = 1e-9, 1e-7


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)).cuda()


def test_c2_full_hot_path_matches_reference(golden):
    """Stages 1–4c at n = 100,000 with the reference's own fitted map: everything the reference stores."""
    from neo_ls_svm_b200 import _lib, _primal

    g = golden("c2_full")
    X, y, _, Xt, _ = load_case("c2_full")
    rows = g["rows"]
    assert np.array_equal(rows, golden_row_subset(len(y)))
    y_ = np.where(y == np.unique(y)[0], -1.0, 1.0)
    s = np.full(len(y), 1.0 / len(y))
    shift, W = g["shift"].ravel(), g["A_map"] / g["scale"].reshape(-1, 1)
    ctx = _lib.Context(0)
    fit = _primal.primal_fit(_dev(X), _dev(y_), _dev(s), _dev(shift), _dev(W), True, ctx=ctx)
    assert fit.opt == int(g["opt"]) and fit.gamma == float(g["gamma"])
    assert rel_err(fit.loo_errors, g["loo_errors"]) < TOL_FIT
    assert_elementwise(fit.loo_errors, g["loo_errors"])
    assert rel_err(fit.beta.cpu().numpy(), g["beta"]) < TOL_FIT
    assert rel_err(fit.beta_eig.cpu().numpy(), g["beta"]) < TOL_FIT
    for key, ref in (("loo_residuals", "loo_residuals"), ("loo_leverage", "loo_leverage"), ("residuals", "residuals"),
                     ("loo_std", "loo_std")):
        got = fit.rows[key].cpu().numpy()[rows]
        assert rel_err(got, g[ref]) < TOL_FIT, key
        assert_elementwise(got, g[ref])
    assert abs(fit.loo_score - float(g["loo_score"])) < TOL_FIT
    assert rel_err(np.diag(fit.U.cpu().numpy()), g["L_diag"]) < TOL_FIT
    yhat, _ = ctx.primal_predict(_dev(Xt), _dev(shift), _dev(W), beta=fit.beta)
    assert rel_err(yhat.cpu().numpy(), g["decision"]) < TOL_PRED


def test_c2_full_public_api_matches_reference(golden):
    """`NeoLSSVM().fit` on the full C2 rows: the supervised pre-pass runs on the device here (n·d ≥ 2²¹), and the
    fitted map, β̂, the LOO vectors and every prediction still equal the reference's."""
    g = golden("c2_full")
    X, y, sw, Xt, _ = load_case("c2_full")
    rows = g["rows"]
    model = NeoLSSVM().fit(X, y)
    aff = model.primal_feature_map_.affine_feature_map
    assert rel_err(aff.shift_, g["shift"]) < 1e-12 and rel_err(aff.scale_, g["scale"]) < 1e-12
    assert rel_err(aff.A_, g["A_map"]) < 1e-10
    assert int(np.argmin(np.abs(model.γs_ - model.γ_))) == int(g["opt"]) and model.γ_ == float(g["gamma"])
    assert rel_err(model.β̂_, g["beta"]) < TOL_FIT
    assert rel_err(model.loo_errors_γs_, g["loo_errors"]) < TOL_FIT
    assert rel_err(model.loo_residuals_[rows], g["loo_residuals"]) < TOL_FIT
    assert rel_err(model.loo_ŷ_[rows], g["loo_yhat"]) < TOL_FIT
    assert rel_err(model.loo_leverage_[rows], g["loo_leverage"]) < TOL_FIT
    assert rel_err(model.loo_std_[rows], g["loo_std"]) < TOL_FIT
    assert abs(model.loo_score_ - float(g["loo_score"])) < TOL_FIT
    assert rel_err(model.ŷ_calib_l1_[:16], g["calib_l1_head"]) < TOL_FIT
    assert rel_err(model.decision_function(Xt), g["decision"]) < TOL_PRED
    assert rel_err(model.predict_std(Xt), g["std"]) < TOL_PRED
    assert np.array_equal(model.predict(Xt), g["predict"])
    assert rel_err(model.predict_proba(Xt), g["proba"]) < TOL_PRED
    assert rel_err(model.predict_quantiles(Xt, quantiles=(0.025, 0.5, 0.975)), g["quantiles_accuracy"]) < 1e-6
    assert rel_err(model.predict_interval(Xt, coverage=0.9), g["interval_90"]) < 1e-6


def test_c5_shaped_fit_matches_oracle():
    """d = 128, num_features = 4096 (m = 4097): Dp/Np padding, the 65-tile projection and whichever eigensolver
    `auto` picks above the Jacobi range, against the chunked CPU oracle; predict_std through U⁻¹."""
    import torch

    from neo_ls_svm_b200 import _lib, _primal
    from oracle import neo_oracle as orc

    n, d, D = 12_000, 128, 4096
    X, y = make_regression_rows(n + 300, d, n_informative=64, noise=200.0)
    Xtr, ytr, Xte = X[:n], y[:n], X[n:]
    fm = OrthogonalRandomFourierFeatures(num_features=D).fit(Xtr[:6000], ytr[:6000], np.ones(6000))
    aff = fm.affine_feature_map
    shift, W = fm.device_weights(d)
    ref = orc.primal_fit_chunked(Xtr, ytr, np.ones(n), aff.shift_, aff.scale_, aff.A_, classifier=False, chunk=4096)
    ctx = _lib.Context(0)
    s = np.full(n, 1.0 / n)
    fit = _primal.primal_fit(_dev(Xtr), _dev(ytr), _dev(s), _dev(shift), _dev(W), False, ctx=ctx)
    assert rel_err(fit.A.cpu().numpy(), ref["A"]) < 1e-12
    assert np.max(np.abs(fit.lam.cpu().numpy() - ref["lam"])) < 1e-12 * ref["lam"][-1]
    assert fit.opt == ref["opt"], "selected γ index must equal the oracle's"
    assert rel_err(fit.loo_errors, ref["loo_errors"]) < TOL_FIT
    assert rel_err(fit.beta.cpu().numpy(), ref["beta"]) < TOL_FIT
    for key in ("loo_residuals", "loo_leverage", "residuals", "loo_std"):
        got = fit.rows[key].cpu().numpy()
        assert rel_err(got, ref[key]) < TOL_FIT, key
        assert_elementwise(got, ref[key])
    phi = orc.feature_map(Xte, aff.shift_, aff.scale_, aff.A_)
    Uinv = ctx.triangular_inverse(fit.U)
    ones = torch.ones(D + 1, dtype=torch.float64, device="cuda")
    yhat, sigma = ctx.primal_predict(_dev(Xte), _dev(shift), _dev(W), beta=fit.beta, B=Uinv, w=ones, want_std=True, b_upper=True)
    assert rel_err(yhat.cpu().numpy(), orc.primal_decision(phi, ref["beta"])) < TOL_PRED
    assert rel_err(sigma.cpu().numpy(), orc.primal_std(phi, ref["L"])) < TOL_PRED


@pytest.mark.parametrize("classifier", [False, True])
def test_dual_n4096_matches_oracle(classifier):
    """The dual solve beyond the fixture sizes: n = 4096 rows (32 row tiles, several γ tiles) against the einsum-free
    CPU restatement, plus the normal equations of the Cholesky re-solve."""
    from oracle import neo_oracle as orc

    n = 4096
    X, y = make_regression_rows(n + 200, 32, n_informative=16, noise=30.0)
    if classifier:
        y = (y > np.median(y)).astype(np.int64)
    model = NeoLSSVM(dual=True).fit(X[:n], y[:n])
    y_ = np.where(y[:n] == model.classes_[0], -1.0, 1.0) if classifier else y[:n].astype(np.float64)
    ref = orc.dual_fit(model.X_, y_, np.ones(n), classifier)
    assert int(np.argmin(np.abs(model.γs_ - model.γ_))) == ref["opt"]
    assert rel_err(model.loo_errors_γs_, ref["loo_errors"]) < TOL_FIT
    assert rel_err(model.α̂_, ref["alpha"]) < TOL_FIT
    assert rel_err(model.loo_residuals_, ref["loo_residuals"]) < TOL_FIT
    assert_elementwise(model.loo_residuals_, ref["loo_residuals"], rtol=1e-8)
    assert rel_err(model.residuals_, ref["residuals"]) < TOL_FIT
    assert rel_err(model.loo_std_, ref["loo_std"]) < 1e-8
    assert abs(model.loo_score_ - ref["loo_score"]) < TOL_FIT
    aff = model.dual_feature_map_
    Xq = orc.affine_map(X[n:], aff.shift_, aff.scale_, aff.A_)
    assert rel_err(model.decision_function(X[n:]), orc.dual_decision(Xq, model.X_, ref["alpha"])) < TOL_PRED
    assert rel_err(model.predict_std(X[n:]), orc.dual_std(Xq, model.X_, ref["L"])) < TOL_PRED
