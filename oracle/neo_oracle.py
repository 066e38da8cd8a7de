"""CPU oracle for the Neo LS-SVM fit/predict hot path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the reference algorithm (lsorber/neo-ls-svm v0.3.0).  It
exists to *check* the CUDA path; it is never imported by the product package `neo_ls_svm_b200`.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import it.

Parity pin: the reference ships no golden vectors / known-answer tests for this path,
so the oracle is pinned against outputs of the reference itself, run in the build container by
`oracle/gen_golden.py` and committed under `tests/golden/*.npz` (see `tests/test_oracle_golden.py`).

Two families of functions:

* ``*_materialised`` follow the reference operation by operation (same temporaries, same BLAS-3 call
  pattern).  They are what the CPU baseline times.
* ``*_chunked`` / ``dual_fit`` are algebraically equivalent restatements with O(chunk) memory, used
  as the oracle for shapes the reference cannot hold in RAM.

Reference line numbers refer to /root/reference/src/neo_ls_svm/.
"""

from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve, eigh

N_GAMMAS_PRIMAL = 1024  # _neo_ls_svm.py:146
N_GAMMAS_DUAL = 128  # _neo_ls_svm.py:270


def gamma_grid(num: int, dtype=np.float64) -> np.ndarray:
    """Log-spaced regularisation grid, _neo_ls_svm.py:146 (primal) and :270 (dual)."""
    return np.logspace(np.log10(1e-6), np.log10(20), num, dtype=dtype)


# ---------------------------------------------------------------------------------------------
# Stage 1: feature map
# ---------------------------------------------------------------------------------------------


def affine_map(X: np.ndarray, shift: np.ndarray, scale: np.ndarray, A: np.ndarray) -> np.ndarray:
    """z = (x - shift) diag(1/scale) A, _affine_feature_map.py:81-89."""
    shift = np.reshape(shift, (-1, X.shape[1]))
    scale = np.reshape(scale, (-1, X.shape[1]))
    W = A / scale.T
    if A.shape[1] < A.shape[0]:
        Z = X @ W - shift @ W
    else:
        Z = (X - shift) @ W
    return Z.astype(X.dtype)


def fourier_map(Z: np.ndarray) -> np.ndarray:
    """phi = [exp(-1j z)/sqrt(D) | 1], _feature_maps.py:194-203."""
    n, D = Z.shape
    phi = np.empty((n, D + 1), dtype=np.complex64 if Z.dtype == np.float32 else np.complex128)
    phi[:, :-1] = np.exp(-1j * Z, dtype=phi.dtype) / np.sqrt(D)
    phi[:, -1] = 1
    return phi


def feature_map(X, shift, scale, A) -> np.ndarray:
    """Stage 1 end to end (a1+a2 of SURVEY.md §8a)."""
    return fourier_map(affine_map(X, shift, scale, A))


# ---------------------------------------------------------------------------------------------
# Shared pieces of the gamma sweep
# ---------------------------------------------------------------------------------------------


def _clip_classifier(res: np.ndarray, y: np.ndarray) -> None:
    """Zero residuals on the correct side of the margin, _neo_ls_svm.py:153-155 / :180-182."""
    if res.ndim == 2:
        res[(y > 0)[:, None] & (res > 0)] = 0
        res[(y < 0)[:, None] & (res < 0)] = 0
    else:
        res[(y > 0) & (res > 0)] = 0
        res[(y < 0) & (res < 0)] = 0


def _objective(s: np.ndarray, loo: np.ndarray, classifier: bool):
    """Per-gamma LOO error and selection objective, _neo_ls_svm.py:158-165."""
    err = s @ np.abs(loo)
    if classifier:
        obj = s @ (np.abs(loo) >= 1) + s @ np.maximum(0, np.abs(loo) - 1) + err
    else:
        obj = err
    return err, obj


def weighted_r2(y, yhat, w) -> float:
    """sklearn.metrics.r2_score with sample weights (used at _neo_ls_svm.py:174)."""
    num = np.sum(w * (y - yhat) ** 2)
    den = np.sum(w * (y - np.average(y, weights=w)) ** 2)
    return float(1.0 - num / den)


def weighted_accuracy(y, yhat, w) -> float:
    """sklearn.metrics.accuracy_score with sample weights (used at _neo_ls_svm.py:172)."""
    return float(np.sum(w * (y == yhat)) / np.sum(w))


# ---------------------------------------------------------------------------------------------
# Primal solve, operation-for-operation (what the reference executes)
# ---------------------------------------------------------------------------------------------


def primal_fit_materialised(phi: np.ndarray, y: np.ndarray, s: np.ndarray, classifier: bool) -> dict:
    """Restatement of NeoLSSVM._optimize_beta_gamma for the identity complexity matrix.

    Follows _neo_ls_svm.py:110-189 step by step, including the three n x m x m complex products
    (:134, :137) and the n-right-hand-side Cholesky solve (:184).
    """
    n, m = phi.shape
    s = s / np.sum(s)  # :110
    Sphi = s[:, None] * phi  # :112
    A = Sphi.conj().T @ Sphi  # :113
    A = (A + A.conj().T) / 2  # :114
    c = np.ones(m, dtype=y.dtype)
    c = c / np.mean(np.abs(c)) / phi.size  # :118
    lam, Q = eigh((1 / c[:, None]) * A)  # :120
    CQ_inv = Q.conj().T * (1 / c[None, :])  # :121
    b = Sphi.conj().T @ (s * y)  # :127
    beta_mat = Q * (CQ_inv @ b)[None, :]  # :129
    phi_beta = np.ascontiguousarray(np.real(phi @ beta_mat))  # :134, :142
    h = np.ascontiguousarray(np.real(Sphi @ Q * (CQ_inv @ Sphi.conj().T).T))  # :137, :143
    gammas = gamma_grid(N_GAMMAS_PRIMAL, y.dtype)  # :146
    rg = 1 / (gammas[None, :] + lam[:, None])  # :147
    with np.errstate(divide="ignore", invalid="ignore"):
        loo = (phi_beta @ rg - y[:, None]) / (1 - h @ rg)  # :149
        yhat_loo = y[:, None] + loo  # :150
    if classifier:
        _clip_classifier(loo, y)
    err, obj = _objective(s, loo, classifier)
    opt = int(np.argmin(obj))  # :159
    out = {
        "s": s,
        "A": A,
        "b": b,
        "lam": lam,
        "gammas": gammas,
        "loo_errors": err,
        "objective": obj,
        "opt": opt,
        "gamma": float(gammas[opt]),
        "loo_residuals": loo[:, opt].copy(),  # :167
        "loo_leverage": h @ rg[:, opt],  # :169
        "loo_error": float(err[opt]),
    }
    out["loo_yhat"] = y + out["loo_residuals"]  # :168
    if classifier:
        out["loo_score"] = weighted_accuracy(y, np.sign(yhat_loo[:, opt]), s)  # :172
    else:
        out["loo_score"] = weighted_r2(y, yhat_loo[:, opt], s)  # :174
    out["beta_eig"] = beta_mat @ rg[:, opt]  # :175
    C = np.diag(c).astype(phi.dtype)
    L = cho_factor(out["gamma"] * C + A)  # :177
    beta = cho_solve(L, b)  # :178
    res = np.real(phi @ beta) - y  # :179
    if classifier:
        _clip_classifier(res, y)
    sigma2 = np.ascontiguousarray(np.real(np.sum(phi * cho_solve(L, phi.conj().T).T, axis=1)))  # :184
    loo_sigma2 = sigma2 + (s * sigma2) ** 2 / (1 - out["loo_leverage"])  # :186
    out.update(beta=beta, L=L, residuals=res, sigma2=sigma2, loo_std=np.sqrt(loo_sigma2))
    return out


# ---------------------------------------------------------------------------------------------
# Primal solve, row-chunked single-T restatement (oracle for shapes the reference cannot hold)
# ---------------------------------------------------------------------------------------------


def primal_fit_chunked(
    X: np.ndarray,
    y: np.ndarray,
    s: np.ndarray,
    shift: np.ndarray,
    scale: np.ndarray,
    A_map: np.ndarray,
    classifier: bool,
    chunk: int = 8192,
) -> dict:
    """Same quantities as `primal_fit_materialised`, O(chunk * m) memory.

    Pass 1 accumulates A = phi^H S^2 phi and b = phi^H S^2 y (:112-114, :127); then
    lam, Q = eigh(A / c) (:120) and v = Q^H b / c (:121, :129); pass 2 forms T = phi Q once per chunk
    and uses P = Re(T * v), H = s^2 |T|^2 / c in place of the three products at :134 / :137 (valid
    for the constant diagonal C that all shipped feature maps produce, _feature_maps.py:129-135);
    pass 3 evaluates the per-row outputs (:167-187) at the selected gamma.
    """
    n = X.shape[0]
    D = A_map.shape[1]
    m = D + 1
    s = s / np.sum(s)
    A = np.zeros((m, m), dtype=np.complex128)
    b = np.zeros(m, dtype=np.complex128)
    for i0 in range(0, n, chunk):
        sl = slice(i0, min(n, i0 + chunk))
        phi = feature_map(X[sl], shift, scale, A_map)
        Sphi = s[sl, None] * phi
        A += Sphi.conj().T @ Sphi
        b += Sphi.conj().T @ (s[sl] * y[sl])
    A = (A + A.conj().T) / 2
    c = 1.0 / (n * m)
    lam, Q = eigh(A / c)
    v = (Q.conj().T @ b) / c
    gammas = gamma_grid(N_GAMMAS_PRIMAL)
    rg = 1 / (gammas[None, :] + lam[:, None])
    err = np.zeros(len(gammas))
    cnt = np.zeros(len(gammas))
    hinge = np.zeros(len(gammas))
    for i0 in range(0, n, chunk):
        sl = slice(i0, min(n, i0 + chunk))
        phi = feature_map(X[sl], shift, scale, A_map)
        T = phi @ Q
        P = np.real(T * v[None, :])
        H = (s[sl, None] ** 2) * (T.real**2 + T.imag**2) / c
        with np.errstate(divide="ignore", invalid="ignore"):
            loo = (P @ rg - y[sl, None]) / (1 - H @ rg)
        if classifier:
            _clip_classifier(loo, y[sl])
        a = np.abs(loo)
        err += s[sl] @ a
        if classifier:
            cnt += s[sl] @ (a >= 1)
            hinge += s[sl] @ np.maximum(0, a - 1)
    obj = cnt + hinge + err if classifier else err
    opt = int(np.argmin(obj))
    gamma = float(gammas[opt])
    beta_eig = Q @ (v * rg[:, opt])
    L = cho_factor(gamma * c * np.eye(m) + A)
    beta = cho_solve(L, b)
    w = 1.0 / (c * (lam + gamma))
    out_vec = {k: np.empty(n) for k in ("loo_residuals", "loo_leverage", "residuals", "sigma2")}
    yhat_loo_raw = np.empty(n)
    for i0 in range(0, n, chunk):
        sl = slice(i0, min(n, i0 + chunk))
        phi = feature_map(X[sl], shift, scale, A_map)
        T = phi @ Q
        sig2 = (T.real**2 + T.imag**2) @ w
        lev = s[sl] ** 2 * sig2
        num = np.real(phi @ beta_eig)
        with np.errstate(divide="ignore", invalid="ignore"):
            loo = (num - y[sl]) / (1 - lev)
        yhat_loo_raw[sl] = y[sl] + loo
        if classifier:
            _clip_classifier(loo, y[sl])
        res = np.real(phi @ beta) - y[sl]
        if classifier:
            _clip_classifier(res, y[sl])
        out_vec["loo_residuals"][sl] = loo
        out_vec["loo_leverage"][sl] = lev
        out_vec["residuals"][sl] = res
        out_vec["sigma2"][sl] = sig2
    out = dict(out_vec)
    out.update(
        s=s, A=A, b=b, lam=lam, gammas=gammas, loo_errors=err, objective=obj, opt=opt, gamma=gamma,
        beta=beta, beta_eig=beta_eig, L=L, loo_error=float(err[opt]),
    )
    out["loo_yhat"] = y + out["loo_residuals"]
    out["loo_std"] = np.sqrt(out["sigma2"] + (s * out["sigma2"]) ** 2 / (1 - out["loo_leverage"]))
    if classifier:
        out["loo_score"] = weighted_accuracy(y, np.sign(yhat_loo_raw), s)
    else:
        out["loo_score"] = weighted_r2(y, yhat_loo_raw, s)
    return out


# ---------------------------------------------------------------------------------------------
# Primal predict / predict_std
# ---------------------------------------------------------------------------------------------


def primal_decision(phi: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """yhat = Re(phi beta), _neo_ls_svm.py:663-665."""
    return np.ascontiguousarray(np.real(phi @ beta))


def primal_std(phi: np.ndarray, L) -> np.ndarray:
    """sigma = sqrt(Re sum phi * cho_solve(L, phi^H)^T), _neo_ls_svm.py:467-469, :477."""
    sigma2 = np.real(np.sum(phi * cho_solve(L, phi.conj().T).T, axis=1))
    return np.sqrt(np.ascontiguousarray(sigma2))


# ---------------------------------------------------------------------------------------------
# Dual solve (einsum-free restatement) and dual predict
# ---------------------------------------------------------------------------------------------


def rbf(Xa: np.ndarray, Xb: np.ndarray) -> np.ndarray:
    """exp(-0.5 ||xa - xb||^2): sklearn rbf_kernel(gamma=0.5) as used at _neo_ls_svm.py:261, :474, :669."""
    d2 = (
        np.sum(Xa * Xa, axis=1)[:, None] - 2.0 * (Xa @ Xb.T) + np.sum(Xb * Xb, axis=1)[None, :]
    )
    np.maximum(d2, 0, out=d2)
    if Xa is Xb:
        np.fill_diagonal(d2, 0.0)
    return np.exp(-0.5 * d2)


def dual_fit(Xt: np.ndarray, y: np.ndarray, s: np.ndarray, classifier: bool) -> dict:
    """Restatement of NeoLSSVM._optimize_alpha_gamma for rho = 1 (_neo_ls_svm.py:252-323).

    The n x G x n tensor of :272-282 is never formed: with SQ = sn*Q, R = 1/(gamma + lam),
    hdiag = (SQ*SQ) R, M = (F0 SQ) * SQ and F0 = F with zero diagonal,
        yhat_loo = -(M R) / hdiag * alpha_loo + F0 alpha_loo.
    """
    n = Xt.shape[0]
    s = s / np.sum(s)  # :252
    sn = s / np.median(np.abs(s))  # :253
    Krbf = rbf(Xt, Xt)
    F = Krbf + 1.0  # :261
    lam, Q = np.linalg.eigh(sn[:, None] * F * sn[None, :])  # :265  (rho = 1 => K = F, :263)
    SQ = sn[:, None] * Q
    qsy = Q.T @ (sn * y)
    alpha_mat = SQ * qsy[None, :]  # :268
    gammas = gamma_grid(N_GAMMAS_DUAL, Xt.dtype)  # :270
    R = 1 / (gammas[None, :] + lam[:, None])
    alpha_loo = alpha_mat @ R  # :285
    hdiag = (SQ * SQ) @ R
    hdiag[hdiag == 0] = np.finfo(Xt.dtype).eps  # :281
    F0 = F.copy()
    np.fill_diagonal(F0, 0)  # :283-284
    M = (F0 @ SQ) * SQ
    yhat_loo = -(M @ R) / hdiag * alpha_loo + F0 @ alpha_loo  # :286
    loo = yhat_loo - y[:, None]  # :287
    if classifier:
        _clip_classifier(loo, y)
    err, obj = _objective(s, loo, classifier)
    opt = int(np.argmin(obj))
    gamma = float(gammas[opt])
    out = {
        "s": s, "sn": sn, "lam": lam, "gammas": gammas, "loo_errors": err, "objective": obj,
        "opt": opt, "gamma": gamma, "loo_residuals": loo[:, opt].copy(),
        "loo_error": float(err[opt]),
    }
    out["loo_yhat"] = y + out["loo_residuals"]
    if classifier:
        out["loo_score"] = weighted_accuracy(y, np.sign(yhat_loo[:, opt]), s)
    else:
        out["loo_score"] = weighted_r2(y, yhat_loo[:, opt], s)
    L = cho_factor(gamma * np.diag(sn**-2) + F)  # :313
    alpha = cho_solve(L, y)  # :314
    res = F @ alpha - y  # :315
    if classifier:
        _clip_classifier(res, y)
    sigma2 = 1.0 - np.sum(Krbf * cho_solve(L, Krbf.T).T, axis=1)  # :321-322
    out.update(alpha=alpha, alpha_eig=alpha_loo[:, opt], L=L, residuals=res, loo_std=np.sqrt(sigma2))
    return out


def dual_decision(Xq: np.ndarray, Xt: np.ndarray, alpha: np.ndarray) -> np.ndarray:
    """yhat = K(x, X) alpha + sum(alpha), _neo_ls_svm.py:668-671."""
    return rbf(Xq, Xt) @ alpha + np.sum(alpha)


def dual_std(Xq: np.ndarray, Xt: np.ndarray, L) -> np.ndarray:
    """sigma = sqrt(1 - sum K * cho_solve(L, K^T)^T), _neo_ls_svm.py:473-477."""
    K = rbf(Xq, Xt)
    return np.sqrt(1.0 - np.sum(K * cho_solve(L, K.T).T, axis=1))


# ---------------------------------------------------------------------------------------------
# Conformal quantile epilogue (batch part of predict_quantiles)
# ---------------------------------------------------------------------------------------------


def isotonic_transform(t: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """sklearn IsotonicRegression(out_of_bounds='clip').transform: clip then linear interpolation."""
    t = np.clip(t, xs[0], xs[-1])
    if len(xs) == 1:
        return np.full_like(t, ys[0])
    hi = np.clip(np.searchsorted(xs, t), 1, len(xs) - 1)
    lo = hi - 1
    slope = (ys[hi] - ys[lo]) / (xs[hi] - xs[lo])
    return slope * (t - xs[lo]) + ys[lo]


def quantile_epilogue(
    yhat: np.ndarray,
    sigma: np.ndarray,
    beta_abs: np.ndarray,
    beta_rel: np.ndarray,
    bias_abs: np.ndarray,
    bias_rel: np.ndarray,
    regressor: bool,
    iso_x: np.ndarray | None = None,
    iso_y: np.ndarray | None = None,
) -> np.ndarray:
    """Per-row part of predict_quantiles, _neo_ls_svm.py:566-600.

    beta_* are the fitted coefficient matrices of the two coherent quantile regressors
    (features x quantiles, last row = intercept; _coherent_linear_quantile_regressor.py:241-255).
    """
    feats = [sigma[:, None]]
    if regressor:
        feats.append(np.abs(yhat)[:, None])
    feats.append(np.ones((len(yhat), 1)))
    Xc = np.hstack(feats)
    d_abs = Xc @ beta_abs + bias_abs[None, :]
    d_rel = np.abs(yhat)[:, None] * (Xc @ beta_rel + bias_rel[None, :])
    both = np.dstack([d_abs, d_rel])  # :578-583
    pick = np.argmin(np.std(both, axis=1), axis=-1)  # :586-588
    delta = both[np.arange(len(yhat)), :, pick]
    q = yhat[:, None] + delta  # :590
    if not regressor:
        q = np.hstack([isotonic_transform(q[:, j], iso_x, iso_y)[:, None] for j in range(q.shape[1])])
        q = np.dstack([1 - q[:, ::-1], q])  # :600
    return q
