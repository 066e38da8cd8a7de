"""Host pre-pass parity: the re-implemented AffineSeparator / ORF fit must reproduce the reference's
fitted shift_, scale_ and folded A_ (= W·scale) for identical inputs and seeds (CPU only)."""

import numpy as np
import pytest

from conftest import rel_err  # noqa: E402
from neo_ls_svm_b200 import AffineSeparator, OrthogonalRandomFourierFeatures
from neo_ls_svm_b200._quantizer import hist_quantized_ecdf, sample_bins_quantized_ecdf
from neo_ls_svm_b200._weighted_quantile import weighted_quantile
from neo_ls_svm_b200.datasets import load_case


def _targets(name, g):
    X, y, sw, _, est = load_case(name)
    if bool(g["classifier"]):
        y_ = np.where(y == np.unique(y)[0], -1.0, 1.0)
    else:
        y_ = y.astype(np.float64)
    s = np.ones(len(y)) if sw is None else sw.astype(np.float64)
    return X, y_, s, est


@pytest.mark.parametrize("name", ["reg_small", "clf_small", "c1", "c2_small", "c3_small"])
def test_orf_fit_matches_reference(name, golden):
    g = golden(name)
    X, y_, s, est = _targets(name, g)
    fm = OrthogonalRandomFourierFeatures(num_features=est.get("num_features", 512)).fit(X, y_, s)
    aff = fm.affine_feature_map
    assert rel_err(aff.shift_, g["shift"]) < 1e-13
    assert rel_err(aff.scale_, g["scale"]) < 1e-13
    assert rel_err(aff.A_, g["A_map"]) < 1e-12
    shift, W = fm.device_weights(X.shape[1])
    assert W.shape == (X.shape[1], fm.D) and shift.shape == (X.shape[1],)
    assert rel_err(W, g["A_map"] / g["scale"].reshape(-1, 1)) < 1e-12


@pytest.mark.parametrize("name", ["dual_reg", "dual_clf"])
def test_separator_fit_matches_reference(name, golden):
    g = golden(name)
    X, y_, s, _ = _targets(name, g)
    sep = AffineSeparator().fit(X, y_, s)
    assert rel_err(sep.shift_, g["shift"]) < 1e-13
    assert rel_err(sep.scale_, g["scale"]) < 1e-13
    assert rel_err(sep.A_, g["A_map"]) < 1e-12
    assert rel_err(sep.transform(X), g["Xt_train"]) < 1e-12


def test_weighted_quantile_toy():
    # The toy example documented in the reference (_weighted_quantile.py:69-70).
    a, w = np.array([0.0, 1.0, 1.0]), np.array([2.0, 1.0, 1.0])
    assert weighted_quantile(a, w, 0.5)[0] == pytest.approx(0.5)
    A = np.stack([a, a[::-1]], axis=1)
    Wt = np.stack([w, w[::-1]], axis=1)
    np.testing.assert_allclose(weighted_quantile(A, Wt, 0.5, axis=0).ravel(), [0.5, 0.5])


def test_quantizer_properties():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(5000)
    hist, edges = hist_quantized_ecdf(x)
    assert hist.sum() == len(x) and len(edges) == len(hist) + 1
    assert np.all(np.diff(edges) > 0) and hist.max() <= 0.125 * len(x) + 1
    # sample_bins quantises the *ranks* of x (reference _quantizer.py:248-252).
    codes = np.unique(x, return_inverse=True)[1]
    hist_r, _ = hist_quantized_ecdf(codes)
    bins = sample_bins_quantized_ecdf(x)
    assert bins.min() == 0 and bins.max() == len(hist_r) - 1
    np.testing.assert_array_equal(np.bincount(bins), hist_r)
    # Few distinct values: the class codes are returned unchanged.
    y = rng.integers(0, 3, size=1000)
    np.testing.assert_array_equal(sample_bins_quantized_ecdf(y), y)
