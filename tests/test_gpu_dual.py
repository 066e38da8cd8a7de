"""Parity of the CUDA dual path (n×n kernel matrix, eigendecomposition, 128-γ LOO sweep, predict) against
the reference's golden outputs and the einsum-free CPU oracle."""

import pickle

import numpy as np
import pytest

from conftest import rel_err  # noqa: E402
from neo_ls_svm_b200 import NeoLSSVM
from neo_ls_svm_b200.datasets import load_case, make_regression_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["dual_reg", "dual_clf"])
def test_dual_estimator_matches_reference(name, golden):
    g = golden(name)
    X, y, sw, Xt, _ = load_case(name)
    model = NeoLSSVM(dual=True).fit(X, y, sample_weight=sw)
    assert model.dual_ and not model.primal_
    assert int(np.argmin(np.abs(model.γs_ - model.γ_))) == int(g["opt"])
    assert rel_err(model.X_, g["Xt_train"]) < 1e-12
    assert rel_err(model.loo_errors_γs_, g["loo_errors"]) < 1e-9
    assert rel_err(model.α̂_, g["alpha"]) < 1e-9
    assert rel_err(model.loo_residuals_, g["loo_residuals"]) < 1e-8
    assert rel_err(model.loo_ŷ_, g["loo_yhat"]) < 1e-8
    assert rel_err(model.residuals_, g["residuals"]) < 1e-9
    assert rel_err(model.loo_std_, g["loo_std"]) < 1e-8
    assert abs(model.loo_score_ - float(g["loo_score"])) < 1e-9
    assert rel_err(np.diag(model.L_[0]), g["L_diag"]) < 1e-9
    assert rel_err(model.decision_function(Xt), g["decision"]) < 1e-7
    assert rel_err(model.predict_std(Xt), g["std"]) < 1e-7
    if bool(g["classifier"]):
        assert np.array_equal(model.predict(Xt), g["predict"])
        assert rel_err(model.predict_proba(Xt), g["proba"]) < 1e-7
    assert rel_err(model.predict_quantiles(Xt, quantiles=(0.025, 0.5, 0.975)), g["quantiles_accuracy"]) < 1e-6
    # device state is a cache: predictions survive a pickle round trip (rebuilt from L_)
    clone = pickle.loads(pickle.dumps(model))
    assert rel_err(clone.decision_function(Xt), g["decision"]) < 1e-7
    assert rel_err(clone.predict_std(Xt), g["std"]) < 1e-7


@pytest.mark.parametrize("n,p", [(257, 5), (1500, 33)])
def test_dual_sweep_matches_oracle(n, p):
    """Shapes the golden fixtures do not cover (ragged n, p not a multiple of the tile K), vs the CPU oracle."""
    import torch

    from neo_ls_svm_b200 import _lib
    from neo_ls_svm_b200._primal import gamma_grid
    from oracle import neo_oracle as orc

    rng = np.random.default_rng(n)
    Xt = rng.standard_normal((n, p)) * 0.6
    y = np.tanh(Xt[:, 0]) + 0.2 * rng.standard_normal(n)
    s = rng.uniform(0.5, 1.5, n)
    ref = orc.dual_fit(Xt, y, s, classifier=False)
    ctx = _lib.Context(0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()  # noqa: E731
    yd, snd = dev(y), dev(ref["sn"])
    sums, yhat_loo, lam = ctx.dual_sweep(dev(Xt), yd, dev(ref["s"]), snd, dev(gamma_grid(128)), False)
    assert rel_err(lam.cpu().numpy(), ref["lam"]) < 1e-11
    assert rel_err(sums[0].cpu().numpy(), ref["loo_errors"]) < 1e-9
    assert int(np.argmin(sums[0].cpu().numpy())) == ref["opt"]
    # the sweep's kernel matrix and eigenbasis are context state: finalize must be handed the sweep's own y / sn
    with pytest.raises(_lib.NlsError):
        ctx.dual_finalize(n, dev(y), dev(ref["sn"]), ref["gamma"])
    fin = ctx.dual_finalize(n, yd, snd, ref["gamma"])
    assert rel_err(fin["alpha"].cpu().numpy(), ref["alpha"]) < 1e-9
    assert rel_err(fin["alpha_eig"].cpu().numpy(), ref["alpha"]) < 1e-8
    assert rel_err(np.sqrt(fin["sigma2"].cpu().numpy()), ref["loo_std"]) < 1e-8
    assert rel_err(yhat_loo[:, ref["opt"]].cpu().numpy() - y, ref["loo_residuals"]) < 1e-8


def test_auto_mode_uses_dual_for_small_n():
    X, y = make_regression_rows(300, 5, n_informative=3, noise=5.0)
    model = NeoLSSVM().fit(X[:250], y[:250])
    assert model.dual_
    assert model.predict(X[250:]).shape == (50,)
    assert model.score(X[250:], y[250:]) > 0.5
