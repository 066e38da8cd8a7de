"""Rows sharded over several GPUs behind the unchanged estimator API (NLS_DEVICES / set_devices).

Needs at least two GPUs (skipped otherwise: the round-end GPU tier runs on one).  The sharded fit must select the same
γ index as the single-GPU fit and agree on every fitted vector to 1e-12 (the partial Grams are summed in a different
order, nothing else changes); the sharded batched predict must equal the single-GPU one bitwise (rows are independent).
"""

import numpy as np
import pytest

from conftest import rel_err  # noqa: E402

pytestmark = pytest.mark.gpu


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("classifier", [False, True])
def test_sharded_fit_equals_single_gpu(classifier):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    import neo_ls_svm_b200 as nls
    from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures
    from neo_ls_svm_b200.datasets import make_churn_rows, make_regression_rows

    if classifier:
        X, y = make_churn_rows(30_000, 20, 8)
    else:
        X, y = make_regression_rows(30_000, 12, n_informative=6, noise=30.0)
    Xtr, ytr, Xte = X[:28_000], y[:28_000], X[28_000:]
    kw = dict(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=256), dual=False)
    nls.set_devices([0])
    one = NeoLSSVM(**kw).fit(Xtr, ytr)
    p1, s1 = one.decision_function(Xte), one.predict_std(Xte)
    nls.set_devices(list(range(min(_gpus(), 4))))
    try:
        many = NeoLSSVM(**kw).fit(Xtr, ytr)
        from neo_ls_svm_b200 import _neo_ls_svm

        _neo_ls_svm._SHARDED_PREDICT_MIN_ROWS = 1  # exercise the sharded predict on the small test batch
        pm, sm = many.decision_function(Xte), many.predict_std(Xte)
    finally:
        nls.set_devices(None)
        _neo_ls_svm._SHARDED_PREDICT_MIN_ROWS = 1 << 16
    assert many.γ_ == one.γ_
    assert rel_err(many.β̂_, one.β̂_) < 1e-10
    assert rel_err(many.loo_errors_γs_, one.loo_errors_γs_) < 1e-12
    assert rel_err(many.loo_residuals_, one.loo_residuals_) < 1e-10
    assert rel_err(many.loo_std_, one.loo_std_) < 1e-10
    assert abs(many.loo_score_ - one.loo_score_) < 1e-12
    assert rel_err(pm, p1) < 1e-10 and rel_err(sm, s1) < 1e-10


def test_device_list_validation():
    import neo_ls_svm_b200 as nls

    if _gpus() < 1:
        pytest.skip("needs a GPU")
    nls.set_devices([0, 0])
    try:
        with pytest.raises(ValueError):
            nls.devices()
        nls.set_devices([99])
        with pytest.raises(ValueError):
            nls.devices()
    finally:
        nls.set_devices(None)
    assert nls.devices() == [0] or len(nls.devices()) >= 1
