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


def test_median_from_crossing_statistics_matches_weighted_quantile():
    """The GPU pre-pass returns only the statistics around the weighted-median crossing; the host formula that
    turns them into the reference's weighted 0.5-quantile is checked here against `weighted_quantile` with a
    NumPy model of the kernel (ties, non-uniform weights, single-element bins)."""
    from neo_ls_svm_b200._binstats import host_median_stats, median_from_stats

    rng = np.random.default_rng(0)
    keys = ("v", "pred", "succ", "w_lt", "w_eq", "n_eq", "w_first", "w_tot")
    checked = 0
    for trial in range(1500):
        n = int(rng.integers(1, 40))
        kind = trial % 4
        if kind == 0:
            a = rng.standard_normal(n)
        elif kind == 1:
            a = rng.integers(0, 4, n).astype(float)  # heavy ties
        elif kind == 2:
            a = np.round(rng.standard_normal(n), 1)
        else:
            a = rng.standard_normal(n)
            a[rng.integers(0, n)] = a[0]
        uniform = trial % 3 != 0
        w = np.full(n, 1.0 / n) if uniform else rng.uniform(0.1, 1, n)
        w = w / w.sum()
        if not uniform and len(np.unique(a)) < n:
            if np.max(np.unique(a, return_counts=True)[1]) > 2:
                continue  # the reference itself is sort-order dependent here
            o = np.argsort(a, kind="stable")
            c = np.cumsum(w[o])
            ref = 0.5 * np.interp(0.5, (c - w[o]) / c[-1], a[o]) + 0.5 * np.interp(0.5, c / c[-1], a[o])
        else:
            ref = weighted_quantile(a, w, 0.5)[0]
        st = host_median_stats(a, w)
        got = median_from_stats(*[np.array([st[k]]) for k in keys])[0]
        assert abs(got - ref) <= 1e-12 * (abs(ref) + np.max(np.abs(a))), (trial, a, w)
        checked += 1
    assert checked > 1000


def test_bin_layout_tiles_cover_each_bin_once():
    from neo_ls_svm_b200._binstats import TILE_ROWS, bin_layout

    rows = [np.arange(0, 5000, 2), np.arange(1, 5000, 2), np.array([5000])]
    s_bins = [np.full((1, len(r)), 1.0 / len(r)) for r in rows]
    perm, w, tiles, bin_tiles = bin_layout(rows, s_bins)
    assert perm.shape == (5001,) and w.shape == (5001,)
    for b, (t0, t1) in enumerate(bin_tiles):
        covered = np.concatenate([np.arange(r0, r1) for (bb, r0, r1, _) in tiles[t0:t1]])
        assert np.all(tiles[t0:t1, 0] == b) and np.all(tiles[t0:t1, 2] - tiles[t0:t1, 1] <= TILE_ROWS)
        np.testing.assert_array_equal(np.sort(perm[covered]), rows[b])


def test_sample_bins_shortcut_equals_the_generic_quantiser():
    """`sample_bins_quantized_ecdf` takes the distinct values/counts of the dense rank codes from a bincount
    instead of sorting a second time; the bins must be those of the generic Quantizer (reference :246-253)."""
    from neo_ls_svm_b200._quantizer import Quantizer, sample_bins_quantized_ecdf, unique_values

    rng = np.random.default_rng(11)
    for y in (rng.standard_normal(5000), np.round(rng.standard_normal(5000), 1), rng.exponential(2.0, 3000)):
        distinct, codes = np.unique(y, return_inverse=True)
        v, inv, cnt = unique_values(y, return_inverse=True, return_counts=True)  # small: the NumPy path
        assert np.array_equal(v, distinct) and np.array_equal(inv, codes) and cnt.sum() == len(y)
        expect = codes if len(distinct) <= np.ceil(np.sqrt(len(codes))) else \
            Quantizer(dtype=np.intp).fit_transform(codes[:, np.newaxis]).ravel()
        assert np.array_equal(sample_bins_quantized_ecdf(y), expect)


def test_weighted_draw_is_randomstate_choice():
    """The validation-free inverse-CDF draw returns RandomState.choice's indices and leaves the same state."""
    from neo_ls_svm_b200._affine import _weighted_draw

    for trial in range(4):
        n = 50_000 + 17 * trial
        w = np.random.default_rng(trial).random(n) + (trial % 2)
        p = w / w.sum()
        r1, r2 = np.random.RandomState(42 + trial), np.random.RandomState(42 + trial)
        a = r1.choice(n, size=1536, p=p)
        b = _weighted_draw(r2, p, 1536)
        assert np.array_equal(a, b) and r1.random_sample() == r2.random_sample()
    # float32 weights (float32 X): `choice` accumulates the CDF in float64, and so must the replica
    n = 200_000
    w32 = (np.random.default_rng(9).random(n) + 0.5).astype(np.float32)
    p32 = w32 / np.sum(w32)
    assert p32.dtype == np.float32
    r1, r2 = np.random.RandomState(7), np.random.RandomState(7)
    a = r1.choice(n, size=1536, p=p32)
    b = _weighted_draw(r2, p32, 1536)
    assert np.array_equal(a, b) and r1.random_sample() == r2.random_sample()


def test_uniform_rank_plan_reproduces_the_reference_median():
    """Uniformly weighted bins (every fit without sample_weight): the ranks and abscissae of `uniform_rank_plan`
    reproduce `weighted_quantile(a, w, 0.5)` from order statistics alone — what the device pre-pass relies on —
    including the bin sizes where the sequential cumulative sum crosses 0.5 by one ulp (n = 26, 28, ...)."""
    from neo_ls_svm_b200._binstats import median_from_ranks, uniform_rank_plan
    from neo_ls_svm_b200._weighted_quantile import weighted_quantile

    rng = np.random.default_rng(0)
    for n_b in list(range(1, 70)) + [100, 101, 1000, 1001, 4096, 65_536, 100_003]:
        w = 1.0 / n_b if n_b % 3 else 1.0 / int(rng.integers(n_b, 5 * n_b + 1))
        a = rng.standard_normal((n_b, 3))
        a[:, 1] = np.round(a[:, 1], 1)  # ties
        ref = weighted_quantile(a, np.full((n_b, 1), w), 0.5, axis=0)[0]
        srt = np.sort(a, axis=0)
        got = median_from_ranks(lambda r: srt[r], n_b, uniform_rank_plan(w, n_b))
        # same ranks, same formula; the reference's numba kernel may contract the last multiply-add
        assert np.max(np.abs(got - ref)) <= 4 * np.finfo(float).eps * max(1.0, np.max(np.abs(ref))), n_b


def test_threaded_shuffle_split_equals_sklearn(monkeypatch):
    """The conformal calibration split of a large fit (own permutation + threaded gathers) returns exactly what
    sklearn.model_selection.train_test_split returns for the same arrays, train size and random state."""
    from sklearn.model_selection import train_test_split

    from neo_ls_svm_b200 import _neo_ls_svm as est

    monkeypatch.setattr(est, "_THREADED_SPLIT_MIN_ROWS", 1000)
    rng = np.random.default_rng(3)
    n = 54_321
    arrays = [rng.standard_normal(n), rng.standard_normal(n).astype(np.float32), rng.standard_normal(n), np.ones(n)]
    for rs in (42, None, np.random.RandomState(7)):
        rs_a, rs_b = (np.random.RandomState(7), np.random.RandomState(7)) if isinstance(rs, np.random.RandomState) else (rs, rs)
        if rs is None:
            np.random.seed(11)
        ref = train_test_split(*arrays, train_size=1440, random_state=rs_a)
        if rs is None:
            np.random.seed(11)
        got = est._shuffle_split(*arrays, train_size=1440, random_state=rs_b)
        assert len(ref) == len(got) == 8
        for a, b in zip(ref, got):
            assert a.dtype == b.dtype and np.array_equal(a, b)
