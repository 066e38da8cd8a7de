"""Parity of the CUDA primal hot path (through the C ABI) against the reference's golden outputs and
the CPU oracle.  Tolerances are the north star's: 1e-9 relative on β̂ / LOO residuals / γ (same index
required), 1e-7 on predictions and quantiles."""

import os

import numpy as np
import pytest

from conftest import assert_elementwise, rel_err  # noqa: E402
from neo_ls_svm_b200.datasets import load_case, make_regression_rows

pytestmark = pytest.mark.gpu

TOL_FIT = 1e-9  # β̂, LOO residuals, LOO error curve
TOL_PRED = 1e-7  # predictions, std, quantiles

PRIMAL = ["reg_small", "clf_small", "c1", "c2_small", "c3_small"]


@pytest.fixture(scope="module")
def gpu():
    import torch

    from neo_ls_svm_b200 import _lib, _primal

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    ctx = _lib.Context(0)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64 if not np.iscomplexobj(a) else np.complex128)).cuda()

    return ctx, dev, _primal, torch


def _case(name, g):
    X, y, sw, Xt, _ = load_case(name)
    classifier = bool(g["classifier"])
    y_ = np.where(y == np.unique(y)[0], -1.0, 1.0) if classifier else y.astype(np.float64)
    s = np.ones(len(y)) if sw is None else sw.astype(np.float64)
    shift = g["shift"].ravel()
    W = g["A_map"] / g["scale"].reshape(-1, 1)
    return X, y_, s / s.sum(), Xt, classifier, shift, W


@pytest.mark.parametrize("name", PRIMAL)
@pytest.mark.parametrize("chunk,stash", [(640, True), (32768, False)])
def test_primal_fit_matches_reference(name, chunk, stash, golden, gpu):
    ctx, dev, _primal, torch = gpu
    g = golden(name)
    X, y_, s, Xt, classifier, shift, W = _case(name, g)
    ctx.set_chunk_rows(chunk)
    fit = _primal.primal_fit(dev(X), dev(y_), dev(s), dev(shift), dev(W), classifier, ctx=ctx, stash=stash)
    assert fit.opt == int(g["opt"]), "selected γ index must equal the reference's"
    assert fit.gamma == float(g["gamma"])
    assert rel_err(fit.loo_errors, g["loo_errors"]) < TOL_FIT
    assert rel_err(fit.beta.cpu().numpy(), g["beta"]) < TOL_FIT
    assert rel_err(fit.beta_eig.cpu().numpy(), g["beta"]) < TOL_FIT
    assert rel_err(fit.rows["loo_residuals"].cpu().numpy(), g["loo_residuals"]) < TOL_FIT
    assert_elementwise(fit.rows["loo_residuals"].cpu().numpy(), g["loo_residuals"])
    assert_elementwise(fit.loo_errors, g["loo_errors"])
    assert rel_err(fit.rows["loo_leverage"].cpu().numpy(), g["loo_leverage"]) < TOL_FIT
    assert rel_err(fit.rows["residuals"].cpu().numpy(), g["residuals"]) < TOL_FIT
    assert rel_err(fit.rows["loo_std"].cpu().numpy(), g["loo_std"]) < TOL_FIT
    assert abs(fit.loo_score - float(g["loo_score"])) < TOL_FIT
    U = fit.U.cpu().numpy()
    probe = np.cos(np.arange(U.shape[0]) * 0.7) + 0.25
    assert rel_err(np.diag(U), g["L_diag"]) < TOL_FIT
    assert rel_err(np.triu(U) @ probe, g["L_probe"]) < TOL_FIT
    # predict / predict_std via the eigenbasis and via the inverse Cholesky factor
    w = _primal.variance_weights(fit.lam, fit.inv_c, fit.gamma)
    yhat, sigma = ctx.primal_predict(dev(Xt), dev(shift), dev(W), beta=fit.beta, B=fit.Q, w=w, want_std=True)
    assert rel_err(yhat.cpu().numpy(), g["decision"]) < TOL_PRED
    assert rel_err(sigma.cpu().numpy(), g["std"]) < TOL_PRED
    m = U.shape[0]
    Uinv = torch.linalg.solve_triangular(torch.triu(fit.U), torch.eye(m, dtype=torch.complex128, device="cuda"), upper=True)
    ones = torch.ones(m, dtype=torch.float64, device="cuda")
    _, sigma2 = ctx.primal_predict(dev(Xt), dev(shift), dev(W), B=Uinv.contiguous(), w=ones, want_std=True)
    assert rel_err(sigma2.cpu().numpy(), g["std"]) < TOL_PRED
    # ... and with the triangular shortcut (only the non-zero half of the contraction), on the raw factor whose
    # strict lower triangle holds garbage as scipy.linalg.cho_factor's does
    Uinv_dirty = Uinv + torch.tril(torch.full_like(Uinv, 7.0), diagonal=-1)
    _, sigma3 = ctx.primal_predict(dev(Xt), dev(shift), dev(W), B=Uinv_dirty.contiguous(), w=ones, want_std=True, b_upper=True)
    assert rel_err(sigma3.cpu().numpy(), g["std"]) < TOL_PRED


@pytest.mark.parametrize("name", ["reg_small", "c3_small"])
def test_feature_map_matches_oracle(name, golden, gpu):
    from oracle import neo_oracle as orc

    ctx, dev, _, _ = gpu
    g = golden(name)
    X, *_ , shift, W = _case(name, g)
    ctx.set_chunk_rows(512)
    phi = ctx.feature_map(dev(X[:1300]), dev(shift), dev(W)).cpu().numpy()
    ref = orc.feature_map(X[:1300], g["shift"], g["scale"], g["A_map"])
    assert rel_err(phi, ref) < 1e-13
    assert rel_err(phi[:16], g["phi_head"]) < 1e-13
    assert np.all(phi[:, -1] == 1.0)


@pytest.mark.parametrize("n,d,D", [(130, 3, 40), (1000, 7, 100), (2049, 5, 129), (777, 9, 64)])
def test_ragged_shapes_match_oracle(n, d, D, gpu):
    """Rows/features/inputs that are not multiples of any tile size, odd d, tail chunks, zero weights."""
    from oracle import neo_oracle as orc

    ctx, dev, _primal, torch = gpu
    rng = np.random.default_rng(n + d + D)
    X = rng.standard_normal((n, d))
    y = np.sin(X[:, 0]) + 0.3 * rng.standard_normal(n)
    s = rng.uniform(0.5, 1.5, n)
    s[:: 17] = 0.0  # zero-weight rows stay in the primal solve (only the dual path drops them)
    shift = rng.standard_normal(d) * 0.1
    scale = rng.uniform(0.5, 2.0, (1, d))
    A_map = rng.standard_normal((d, D)) * 0.7
    ref = orc.primal_fit_chunked(X, y, s, shift, scale, A_map, classifier=False, chunk=500)
    ctx.set_chunk_rows(256)
    W = A_map / scale.T
    fit = _primal.primal_fit(dev(X), dev(y), dev(s / s.sum()), dev(shift), dev(W), False, ctx=ctx)
    assert rel_err(fit.A.cpu().numpy(), ref["A"]) < 1e-12
    assert rel_err(fit.b.cpu().numpy(), ref["b"]) < 1e-12
    assert rel_err(fit.lam.cpu().numpy(), ref["lam"]) < 1e-11
    assert fit.opt == ref["opt"]
    assert rel_err(fit.loo_errors, ref["loo_errors"]) < TOL_FIT
    assert rel_err(fit.beta.cpu().numpy(), ref["beta"]) < TOL_FIT
    assert rel_err(fit.rows["loo_residuals"].cpu().numpy(), ref["loo_residuals"]) < TOL_FIT
    # These tiny problems (n = 130, m = 41, features of 3 inputs) have eigenvalues down to rounding level next to the
    # selected γ: two correctly rounded Gram matrices (FP64 DMMA summation vs the exact INT8 summation, which
    # test_int8_gram_is_as_accurate_as_fp64_summation shows to be the closer one) move single LOO residuals by 1e-11 of
    # the largest, so the absolute part of the elementwise bound is 5e-11 here; the norm-wise 1e-9 above is unchanged.
    assert_elementwise(fit.rows["loo_residuals"].cpu().numpy(), ref["loo_residuals"], atol_scale=5e-11)
    assert rel_err(fit.rows["loo_std"].cpu().numpy(), ref["loo_std"]) < TOL_FIT


def test_tma_pipeline_equals_plain_loader(golden, gpu):
    """The TMA/mbarrier pipeline and the bounds-checked reference loader must agree bitwise."""
    from neo_ls_svm_b200 import _lib

    ctx, dev, _primal, _ = gpu
    g = golden("clf_small")
    X, y_, s, Xt, classifier, shift, W = _case("clf_small", g)
    os.environ["NLS_NO_TMA"] = "1"
    try:
        plain = _lib.Context(0)
    finally:
        os.environ["NLS_NO_TMA"] = "0"
    ctx.set_chunk_rows(512)
    plain.set_chunk_rows(512)
    A1, b1 = ctx.primal_gram(dev(X), dev(y_), dev(s), dev(shift), dev(W))
    A2, b2 = plain.primal_gram(dev(X), dev(y_), dev(s), dev(shift), dev(W))
    assert np.array_equal(A1.cpu().numpy(), A2.cpu().numpy())
    assert np.array_equal(b1.cpu().numpy(), b2.cpu().numpy())


def test_determinism(golden, gpu):
    """Fixed-order reductions: two runs are bitwise identical (the reference is)."""
    ctx, dev, _primal, _ = gpu
    g = golden("reg_small")
    X, y_, s, Xt, classifier, shift, W = _case("reg_small", g)
    ctx.set_chunk_rows(512)
    f1 = _primal.primal_fit(dev(X), dev(y_), dev(s), dev(shift), dev(W), classifier, ctx=ctx)
    f2 = _primal.primal_fit(dev(X), dev(y_), dev(s), dev(shift), dev(W), classifier, ctx=ctx)
    assert np.array_equal(f1.loo_errors, f2.loo_errors)
    assert np.array_equal(f1.beta.cpu().numpy(), f2.beta.cpu().numpy())
    assert np.array_equal(f1.rows["loo_residuals"].cpu().numpy(), f2.rows["loo_residuals"].cpu().numpy())


def test_large_n_properties(gpu):
    """At a size the reference cannot hold (n = 300k, m = 1025): invariants that do not need an oracle."""
    ctx, dev, _primal, torch = gpu
    n, d, D = 300_000, 16, 1024
    X, y = make_regression_rows(n, d, n_informative=8, noise=5.0)
    rng = np.random.default_rng(0)
    W = rng.standard_normal((d, D)) * 0.5
    shift = np.zeros(d)
    s = np.full(n, 1.0 / n)
    ctx.set_chunk_rows(32768)
    Xd, yd, sd, shd, Wd = dev(X), dev(y), dev(s), dev(shift), dev(W)
    fit = _primal.primal_fit(Xd, yd, sd, shd, Wd, False, ctx=ctx)
    A = fit.A
    assert torch.allclose(A, A.conj().T, rtol=0, atol=0), "A must be exactly Hermitian"
    # trace(A) = Σ s_i² ||φ_i||² = Σ s_i² (1 + 1) since ||exp(-iz)/√D||² = 1 and the constant feature is 1
    assert abs(float(A.diagonal().real.sum()) - 2.0 * float((sd * sd).sum())) < 1e-12 * 2.0 / n
    # eigen-expansion β̂ equals the Cholesky re-solve, and both solve the normal equations
    assert rel_err(fit.beta_eig.cpu().numpy(), fit.beta.cpu().numpy()) < 1e-8
    M = A + (fit.gamma / fit.inv_c) * torch.eye(D + 1, dtype=A.dtype, device=A.device)
    assert rel_err((M @ fit.beta).cpu().numpy(), fit.b.cpu().numpy()) < 1e-9
    # Σ s_i |loo_i| recomputed from the per-row output equals the swept error at the optimum
    err = float((sd * fit.rows["loo_residuals"].abs()).sum())
    assert abs(err - fit.loo_error) < 1e-9 * abs(fit.loo_error)
    # linearity in y: scaling y scales β̂ and the residuals, and keeps γ
    fit2 = _primal.primal_fit(Xd, 3.0 * yd, sd, shd, Wd, False, ctx=ctx)
    assert fit2.opt == fit.opt
    assert rel_err(fit2.beta.cpu().numpy(), 3.0 * fit.beta.cpu().numpy()) < 1e-9
    # sharding invariance: two half-shards summed == one shard (what the NCCL all-reduce does)
    h = n // 2
    A1, b1 = ctx.primal_gram(Xd[:h].contiguous(), yd[:h].contiguous(), sd[:h].contiguous(), shd, Wd)
    A2, b2 = ctx.primal_gram(Xd[h:].contiguous(), yd[h:].contiguous(), sd[h:].contiguous(), shd, Wd)
    assert rel_err((A1 + A2).cpu().numpy(), A.cpu().numpy()) < 1e-12


def test_quantile_epilogue_matches_reference(golden, gpu):
    ctx, dev, _, _ = gpu
    for name in ("reg_small", "clf_small", "c1"):
        g = golden(name)
        classifier = bool(g["classifier"])
        q = ctx.quantile_epilogue(
            dev(g["decision"]), dev(g["std"]), dev(g["cqr_abs_beta"]), dev(g["cqr_rel_beta"]), dev(g["cqr_abs_bias"]),
            dev(g["cqr_rel_bias"]), not classifier,
            dev(g["iso_x"]) if classifier else None, dev(g["iso_y"]) if classifier else None)
        assert q.shape == g["quantiles_accuracy"].shape
        assert rel_err(q.cpu().numpy(), g["quantiles_accuracy"]) < TOL_PRED


def test_error_paths(gpu):
    from neo_ls_svm_b200 import _lib

    ctx, dev, _, torch = gpu
    X = dev(np.zeros((4, 2)))
    with pytest.raises(_lib.NlsError):
        ctx.lib.nls_ctx_set_chunk_rows(ctx.handle, 5) and _lib.check(-1)
    with pytest.raises(_lib.NlsError):
        _lib.check(ctx.lib.nls_feature_map(ctx.handle, None, 4, 2, None, None, 8, None))
    assert b"null" in ctx.lib.nls_last_error()
    with pytest.raises(_lib.NlsError):
        _lib.check(ctx.lib.nls_ctx_set_gemm_core(ctx.handle, 7))
    with pytest.raises(KeyError):
        ctx.set_gemm_core("fp8")
    del X


def test_int8_sweep_variants_agree(golden, gpu):
    """The three INT8 routes of the sweep — projection epilogue writing the sweep's operand planes (spill column at a k-step
    boundary: m = 1025), planes sliced from FP64 P / U with row exponents (m = 130: no spill), and the DMMA sweep behind the
    INT8 projection — give the same LOO curve and σ² to FP64 rounding."""
    from neo_ls_svm_b200 import _lib

    _, dev, _primal, _ = gpu
    g = golden("c3_small")
    X, y_, s, Xt, classifier, shift, W = _case("c3_small", g)
    fits = {}
    for core in ("ozaki", "ozaki-dmma-sweep"):
        ctx = _lib.Context(0)
        ctx.set_gemm_core(core)
        ctx.set_chunk_rows(2048)
        fits[core] = _primal.primal_fit(dev(X), dev(y_), dev(s), dev(shift), dev(W), classifier, ctx=ctx)
        assert fits[core].opt == int(g["opt"])
    a, b = fits["ozaki"], fits["ozaki-dmma-sweep"]
    assert rel_err(a.loo_errors, b.loo_errors) < 1e-12
    assert rel_err(a.rows["loo_std"].cpu().numpy(), b.rows["loo_std"].cpu().numpy()) < 1e-12
    assert rel_err(a.rows["loo_leverage"].cpu().numpy(), b.rows["loo_leverage"].cpu().numpy()) < 1e-12
    # a width whose column count is no multiple of 32 and has no spill column takes the unfused slicing route
    rng = np.random.default_rng(3)
    Xr = rng.standard_normal((3000, 5))
    yr = np.sin(Xr[:, 0]) + 0.1 * rng.standard_normal(3000)
    sr = np.full(3000, 1.0 / 3000)
    Wr = rng.standard_normal((5, 129)) * 0.5
    out = {}
    for core in ("ozaki", "dmma"):
        ctx = _lib.Context(0)
        ctx.set_gemm_core(core)
        ctx.set_chunk_rows(1024)
        out[core] = _primal.primal_fit(dev(Xr), dev(yr), dev(sr), dev(np.zeros(5)), dev(Wr), False, ctx=ctx)
    assert out["ozaki"].opt == out["dmma"].opt
    assert rel_err(out["ozaki"].loo_errors, out["dmma"].loo_errors) < 1e-11
    assert rel_err(out["ozaki"].rows["loo_residuals"].cpu().numpy(), out["dmma"].rows["loo_residuals"].cpu().numpy()) < 1e-10


@pytest.mark.parametrize("kind", ["dc", "jacobi", "cusolver"])
@pytest.mark.parametrize("m", [5, 64, 257, 513, 1025, 2049])
def test_heev_matches_lapack(kind, m, gpu):
    """Stage 3: the hand-written tridiagonalisation + divide-and-conquer solver (default), the hand-written
    block-Jacobi kernels and the cuSOLVER comparator all reproduce LAPACK's spectrum of a Gram-like Hermitian matrix
    graded over 14 decades, with a unitary, residual-free basis."""
    if kind == "jacobi" and m > 1100:
        pytest.skip("the Jacobi comparator is O(sweeps m^3): covered up to m = 1025")
    from neo_ls_svm_b200 import _lib

    _, dev, _, torch = gpu
    ctx = _lib.Context(0)
    ctx.set_eigensolver(kind)
    rng = np.random.default_rng(m)
    U, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    lam_true = np.logspace(2, -12, m)
    A = (U * lam_true) @ U.conj().T
    A = (A + A.conj().T) / 2
    scale = 3.0
    lam, Q = ctx.heev(dev(A), scale)
    lam, Q = lam.cpu().numpy(), Q.cpu().numpy()
    ref = np.linalg.eigvalsh(A * scale)
    assert np.all(np.diff(lam) >= 0), "eigenvalues must be ascending like scipy.linalg.eigh"
    # the Jacobi comparator accumulates ~12 sweeps of rotations: 6e-13 at m = 1025; the default solver holds 1e-14
    assert np.max(np.abs(lam - ref)) < (2e-12 if kind == "jacobi" else 1e-13) * ref[-1]
    assert np.max(np.abs(Q.conj().T @ Q - np.eye(m))) < (1e-12 if kind == "jacobi" else 1e-13)
    assert np.max(np.abs((A * scale) @ Q - Q * lam[None, :])) < 1e-12 * ref[-1]
    if kind == "jacobi":
        assert 1 <= ctx.last_eig_sweeps() <= 40


@pytest.mark.parametrize("m", [1, 2, 3, 31, 32, 33, 34, 63, 65, 100, 129, 300, 777])
def test_dc_eigensolver_odd_sizes_and_clusters(m, gpu):
    """The default eigensolver around its panel (32), leaf (32) and tile boundaries, with 4-fold degenerate
    eigenvalues (even m), a spectrum graded over 18 decades (odd m), and a diagonal / all-zero input."""
    from neo_ls_svm_b200 import _lib

    _, dev, _, torch = gpu
    ctx = _lib.Context(0)
    ctx.set_eigensolver("dc")
    rng = np.random.default_rng(2000 + m)
    U, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    lam_true = np.logspace(0, -18, m) if m % 2 else np.repeat(rng.standard_normal((m + 3) // 4), 4)[:m]
    A = (U * lam_true) @ U.conj().T
    A = (A + A.conj().T) / 2
    for mat in (A, np.diag(lam_true).astype(np.complex128), np.zeros((m, m), dtype=np.complex128)):
        lam, Q = ctx.heev(dev(mat), 1.0)
        lam, Q = lam.cpu().numpy(), Q.cpu().numpy()
        ref = np.linalg.eigvalsh(mat)
        nrm = max(np.abs(ref).max(), 1e-300)
        assert np.all(np.diff(lam) >= 0)
        assert np.max(np.abs(lam - ref)) < 5e-13 * nrm
        assert np.max(np.abs(Q.conj().T @ Q - np.eye(m))) < 1e-12
        assert np.max(np.abs(mat @ Q - Q * lam[None, :])) < 1e-12 * nrm


def test_dc_eigensolver_is_deterministic(gpu):
    """Fixed-order reductions throughout: two solves give bitwise identical eigenpairs."""
    from neo_ls_svm_b200 import _lib

    _, dev, _, torch = gpu
    ctx = _lib.Context(0)
    ctx.set_eigensolver("dc")
    rng = np.random.default_rng(5)
    M = rng.standard_normal((400, 300)) + 1j * rng.standard_normal((400, 300))
    A = dev(M.conj().T @ M / 400)
    l1, Q1 = ctx.heev(A, 2.0)
    l2, Q2 = ctx.heev(A, 2.0)
    assert torch.equal(l1, l2) and torch.equal(Q1, Q2)


@pytest.mark.parametrize("m", [1, 2, 63, 64, 65, 130, 513, 1025])
def test_cholesky_and_triangular_inverse_match_lapack(m, gpu):
    """The hand-written blocked Cholesky (M = UᴴU in scipy's cho_factor layout), its two triangular solves and the
    blocked triangular inverse against LAPACK, on a Hermitian positive definite matrix with condition 1e10; the strict
    lower triangle of the factor keeps garbage and must be ignored by the inverse."""
    import scipy.linalg as sl

    from neo_ls_svm_b200 import _lib

    _, dev, _, torch = gpu
    ctx = _lib.Context(0)
    rng = np.random.default_rng(3000 + m)
    Qm, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    A = (Qm * np.logspace(0, -10, m)) @ Qm.conj().T
    A = (A + A.conj().T) / 2
    b = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    shift = 1e-3
    U, beta = ctx.cholesky_solve(dev(A), shift, dev(b))
    Uh, bh = U.cpu().numpy(), beta.cpu().numpy()
    ref_c, _ = sl.cho_factor(A + shift * np.eye(m), lower=False)
    ref_beta = sl.cho_solve((ref_c, False), b)
    assert rel_err(np.triu(Uh), np.triu(ref_c)) < 1e-12
    assert rel_err(bh, ref_beta) < 1e-10
    dirty = U + torch.tril(torch.full_like(U, 7.0 - 3.0j), diagonal=-1)
    Binv = ctx.triangular_inverse(dirty.contiguous()).cpu().numpy()
    assert np.max(np.abs(np.tril(Binv, -1))) == 0.0
    assert rel_err(Binv, np.linalg.inv(np.triu(ref_c))) < 1e-10
    # real symmetric instantiation (the dual path's factor)
    Ar = np.real(A) + shift * np.eye(m)
    Ur = np.triu(np.linalg.cholesky(Ar).T)
    Br = ctx.triangular_inverse(dev(Ur)).cpu().numpy()
    assert rel_err(Br, np.linalg.inv(Ur)) < 1e-10


@pytest.mark.parametrize("jb", ["8", "4"])
@pytest.mark.parametrize("m", [1, 2, 3, 15, 16, 17, 33, 100, 300])
def test_jacobi_odd_sizes_and_clusters(jb, m, gpu, monkeypatch):
    """Both block widths of the hand-written Jacobi solver (8: ping-pong rounds with 16 x 16 pivots, the default;
    4: the original 8 x 8 pivots) on sizes around the block/pivot boundaries, with 4-fold degenerate eigenvalues
    (even m) or a spectrum graded over 18 decades (odd m): zero padding, single-pivot and clustered cases."""
    from neo_ls_svm_b200 import _lib

    _, dev, _, torch = gpu
    monkeypatch.setenv("NLS_JACOBI_JB", jb)
    ctx = _lib.Context(0)
    ctx.set_eigensolver("jacobi")
    rng = np.random.default_rng(1000 + m)
    U, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    lam_true = np.logspace(0, -18, m) if m % 2 else np.repeat(rng.standard_normal((m + 3) // 4), 4)[:m]
    A = (U * lam_true) @ U.conj().T
    A = (A + A.conj().T) / 2
    lam, Q = ctx.heev(dev(A), 1.0)
    lam, Q = lam.cpu().numpy(), Q.cpu().numpy()
    ref = np.linalg.eigvalsh(A)
    nrm = np.abs(ref).max()
    assert np.all(np.diff(lam) >= 0)
    assert np.max(np.abs(lam - ref)) < 5e-13 * nrm
    assert np.max(np.abs(Q.conj().T @ Q - np.eye(m))) < 1e-12
    assert np.max(np.abs(A @ Q - Q * lam[None, :])) < 1e-12 * nrm


def test_fit_is_eigensolver_independent(golden, gpu):
    """β̂, the LOO error curve and the selected γ do not depend on which eigensolver produced (λ, Q)."""
    from neo_ls_svm_b200 import _lib

    _, dev, _primal, _ = gpu
    g = golden("c1")
    X, y_, s, Xt, classifier, shift, W = _case("c1", g)
    out = {}
    for kind in ("dc", "jacobi", "cusolver"):
        ctx = _lib.Context(0)
        ctx.set_eigensolver(kind)
        out[kind] = _primal.primal_fit(dev(X), dev(y_), dev(s), dev(shift), dev(W), classifier, ctx=ctx)
        assert out[kind].opt == int(g["opt"])
        assert rel_err(out[kind].beta_eig.cpu().numpy(), g["beta"]) < TOL_FIT
        assert rel_err(out[kind].loo_errors, g["loo_errors"]) < TOL_FIT
    assert rel_err(out["jacobi"].lam.cpu().numpy(), out["cusolver"].lam.cpu().numpy()) < 1e-12
    assert rel_err(out["dc"].lam.cpu().numpy(), out["cusolver"].lam.cpu().numpy()) < 1e-12


def test_gram_from_host_rows_equals_device_rows(golden, gpu):
    """nls_primal_gram_h2d (rows streamed from pinned host memory underneath the Gram pass, several row groups and a
    ragged tail) gives bitwise the same A, b as the device-resident entry point, and the uploaded copies are exact."""
    from neo_ls_svm_b200 import _lib

    _, dev, _primal, torch = gpu
    g = golden("c1")
    X, y_, s, Xt, classifier, shift, W = _case("c1", g)
    ctx = _lib.Context(0)
    ctx.set_chunk_rows(512)  # 10,000 rows -> 5 upload groups of 2048 rows, the last one ragged
    A0, b0 = ctx.primal_gram(dev(X), dev(y_), dev(s), dev(shift), dev(W))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()  # noqa: E731
    A1, b1, Xd, yd, sd = ctx.primal_gram_h2d(pin(X), pin(y_), pin(s), dev(shift), dev(W))
    torch.cuda.synchronize()
    assert torch.equal(A0, A1) and torch.equal(b0, b1)
    assert np.array_equal(Xd.cpu().numpy(), X) and np.array_equal(yd.cpu().numpy(), y_) and np.array_equal(sd.cpu().numpy(), s)
    fit = _primal.primal_fit(None, None, None, dev(shift), dev(W), classifier, ctx=ctx, host_rows=(pin(X), pin(y_), pin(s)))
    assert fit.opt == int(g["opt"])
    assert rel_err(fit.beta_eig.cpu().numpy(), g["beta"]) < TOL_FIT


@pytest.mark.parametrize("name", ["c1", "clf_small", "c3_small"])
def test_int8_core_matches_dmma(name, golden, gpu):
    """The Gram and the projection T = φQ on the INT8 tensor cores (tcgen05 kind::i8, Ozaki scheme: 7 digit planes, 28
    exact plane products, FP64 recombination; the default) against the FP64 DMMA core: same γ index, A / LOO curve /
    residuals / std far inside the 1e-9 bar, and both against the reference's golden outputs."""
    from neo_ls_svm_b200 import _lib

    _, dev, _primal, _ = gpu
    g = golden(name)
    X, y_, s, Xt, classifier, shift, W = _case(name, g)
    out = {}
    for core in ("dmma", "ozaki"):
        ctx = _lib.Context(0)
        ctx.set_gemm_core(core)
        ctx.set_chunk_rows(1024)  # several chunks with a ragged tail
        n0 = ctx.launch_count()
        out[core] = fit = _primal.primal_fit(dev(X), dev(y_), dev(s), dev(shift), dev(W), classifier, ctx=ctx)
        out[core + "_launches"] = ctx.launch_count() - n0
        assert fit.opt == int(g["opt"])
        assert rel_err(fit.loo_errors, g["loo_errors"]) < TOL_FIT
        assert rel_err(fit.rows["loo_residuals"].cpu().numpy(), g["loo_residuals"]) < TOL_FIT
        assert_elementwise(fit.rows["loo_residuals"].cpu().numpy(), g["loo_residuals"])
        assert rel_err(fit.rows["loo_std"].cpu().numpy(), g["loo_std"]) < TOL_FIT
    a, b = out["dmma"], out["ozaki"]
    assert out["ozaki_launches"] > out["dmma_launches"], "the INT8 core adds its slicing kernels: it must have run"
    assert rel_err(b.A.cpu().numpy(), a.A.cpu().numpy()) < 1e-13
    assert rel_err(b.beta.cpu().numpy(), a.beta.cpu().numpy()) < 1e-10
    assert rel_err(b.loo_errors, a.loo_errors) < 1e-11
    assert rel_err(b.rows["loo_residuals"].cpu().numpy(), a.rows["loo_residuals"].cpu().numpy()) < 1e-10
    assert rel_err(b.rows["loo_std"].cpu().numpy(), a.rows["loo_std"].cpu().numpy()) < 1e-11


@pytest.mark.parametrize("n,d,D,weights", [(130, 3, 40, "ragged"), (1500, 5, 129, "ragged"), (3000, 6, 256, "uniform")])
def test_int8_gram_is_as_accurate_as_fp64_summation(n, d, D, weights, gpu):
    """The INT8 Gram rounds every operand entry once at 2^-56 of the chunk's largest entry and sums exactly, so its A must
    be at least as close to the exactly summed Gram of the SAME device features (long double on the host) as the FP64
    DMMA Gram is — this is what makes it a drop-in: it moves A by less than FP64 rounding already does."""
    from neo_ls_svm_b200 import _lib

    _, dev, _, _ = gpu
    rng = np.random.default_rng(7 * n + D)
    X = rng.standard_normal((n, d))
    s = np.ones(n) if weights == "uniform" else rng.uniform(0.5, 1.5, n)
    if weights == "ragged":
        s[::17] = 0.0
    s = s / s.sum()
    y = np.sin(X[:, 0])
    shift = rng.standard_normal(d) * 0.1
    W = rng.standard_normal((d, D)) * 0.7
    err = {}
    for core in ("dmma", "ozaki"):
        ctx = _lib.Context(0)
        ctx.set_gemm_core(core)
        ctx.set_chunk_rows(1024)
        A, _ = ctx.primal_gram(dev(X), dev(y), dev(s), dev(shift), dev(W))
        phi = ctx.feature_map(dev(X), dev(shift), dev(W)).cpu().numpy()
        Pw = (s[:, None] * phi).astype(np.clongdouble)
        exact = Pw.conj().T @ Pw  # long double accumulation of the device's own features
        err[core] = float(np.max(np.abs(A.cpu().numpy().astype(np.clongdouble) - exact)) / np.max(np.abs(exact)))
    assert err["ozaki"] < 5e-16
    assert err["ozaki"] <= err["dmma"] + 1e-16, err
