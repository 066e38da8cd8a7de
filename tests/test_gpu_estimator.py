"""End-to-end parity of the public estimator (`NeoLSSVM`) against the reference's golden outputs."""

import pickle

import numpy as np
import pytest

from conftest import rel_err  # noqa: E402
from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures
from neo_ls_svm_b200.datasets import load_case

pytestmark = pytest.mark.gpu


def _fit(name):
    X, y, sw, Xt, est = load_case(name)
    kw = {}
    if "num_features" in est:
        kw["primal_feature_map"] = OrthogonalRandomFourierFeatures(num_features=est["num_features"])
    if "dual" in est:
        kw["dual"] = est["dual"]
    return NeoLSSVM(**kw).fit(X, y, sample_weight=sw), Xt


@pytest.mark.parametrize("name", ["reg_small", "clf_small", "c1", "c2_small"])
def test_estimator_matches_reference(name, golden):
    g = golden(name)
    model, Xt = _fit(name)
    assert int(np.argmin(np.abs(model.γs_ - model.γ_))) == int(g["opt"])
    assert model.γ_ == float(g["gamma"])
    assert rel_err(model.β̂_, g["beta"]) < 1e-9
    assert rel_err(model.loo_errors_γs_, g["loo_errors"]) < 1e-9
    assert rel_err(model.loo_residuals_, g["loo_residuals"]) < 1e-9
    assert rel_err(model.loo_ŷ_, g["loo_yhat"]) < 1e-9
    assert rel_err(model.loo_leverage_, g["loo_leverage"]) < 1e-9
    assert rel_err(model.residuals_, g["residuals"]) < 1e-9
    assert rel_err(model.loo_std_, g["loo_std"]) < 1e-9
    assert abs(model.loo_score_ - float(g["loo_score"])) < 1e-9
    assert rel_err(np.diag(model.L_[0]), g["L_diag"]) < 1e-9
    assert rel_err(model.ŷ_calib_l1_[:16], g["calib_l1_head"]) < 1e-9
    assert rel_err(model.decision_function(Xt), g["decision"]) < 1e-7
    assert rel_err(model.predict_std(Xt), g["std"]) < 1e-7
    if bool(g["classifier"]):
        assert np.array_equal(model.predict(Xt), g["predict"])
        assert rel_err(model.predict_proba(Xt), g["proba"]) < 1e-7
    else:
        assert rel_err(model.predict(Xt), g["predict"]) < 1e-7
    assert rel_err(model.predict_quantiles(Xt, quantiles=(0.025, 0.5, 0.975)), g["quantiles_accuracy"]) < 1e-6
    assert rel_err(model.predict_interval(Xt, coverage=0.9), g["interval_90"]) < 1e-6


def test_pickle_roundtrip_rebuilds_device_state(golden):
    g = golden("reg_small")
    model, Xt = _fit("reg_small")
    clone = pickle.loads(pickle.dumps(model))
    assert "_device_state" not in clone.__dict__
    assert rel_err(clone.decision_function(Xt), g["decision"]) < 1e-7
    assert rel_err(clone.predict_std(Xt), g["std"]) < 1e-7  # via U⁻¹ rebuilt from L_


def test_pandas_passthrough():
    import pandas as pd

    X, y, sw, Xt, _ = load_case("reg_small")
    cols = [f"f{j}" for j in range(X.shape[1])]
    model = NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=64), dual=False)
    model.fit(pd.DataFrame(X, columns=cols), pd.Series(y))
    Xt_df = pd.DataFrame(Xt, columns=cols, index=np.arange(len(Xt)) + 100)
    for method in ("decision_function", "predict", "predict_std", "predict_proba"):
        out_df, out_np = getattr(model, method)(Xt_df), getattr(model, method)(Xt)
        assert isinstance(out_df, pd.Series) and out_df.index.equals(Xt_df.index)
        assert np.array_equal(out_df.to_numpy(), out_np)
    q_df, q_np = model.predict_quantiles(Xt_df), model.predict_quantiles(Xt)
    assert isinstance(q_df, pd.DataFrame) and np.array_equal(q_df.to_numpy(), q_np)
    assert np.all(np.diff(q_np, axis=1) >= -1e-9), "quantiles must be monotone"


def test_float32_inputs_keep_dtype():
    X, y, sw, Xt, _ = load_case("reg_small")
    model = NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=64), dual=False)
    model.fit(X.astype(np.float32), y.astype(np.float32))
    assert model.β̂_.dtype == np.complex64 and model.loo_residuals_.dtype == np.float32
    assert model.predict(Xt.astype(np.float32)).dtype == np.float32
