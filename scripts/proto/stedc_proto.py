"""NumPy prototype of the divide-and-conquer tridiagonal eigensolver (design study for csrc/stedc.cuh).

Same structure as the device version: Jacobi leaves, per merge: host deflation (LAPACK dlaed2 logic), secular roots with
differences relative to the nearest pole, Gu-Eisenstat z-hat, normalised eigenvector coefficients, GEMM."""
import numpy as np

EPS = np.finfo(float).eps


def leaf_eig(d, e):
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    return np.linalg.eigh(T)


def secular_root(i, d, z2, rho):
    """Root i of 1 + rho sum z2_j/(d_j - lam) in (d_i, d_{i+1}) (last: (d_K-1, d_K-1 + rho ||z||^2)).
    Returns (origin index, mu) with lam = d[origin] + mu."""
    K = len(d)
    if i < K - 1:
        left, right = d[i], d[i + 1]
        mid = 0.5 * (right - left)
        # f at the midpoint decides the nearer pole
        dd = d - left
        fmid = 1.0 + rho * np.sum(z2 / (dd - mid))
        if fmid > 0:   # root in the left half: origin d_i
            org = i; lo, hi = 0.0, mid
        else:
            org = i + 1; lo, hi = -mid, 0.0
    else:
        org = K - 1
        lo, hi = 0.0, rho * np.sum(z2)
    dd = d - d[org]          # poles relative to the origin (exact differences of the inputs)
    mu = 0.5 * (lo + hi) if i < K - 1 else min(hi, max(lo, 0.5 * hi))
    for it in range(100):
        delta = dd - mu
        t = z2 / delta
        psi = rho * np.sum(t[: i + 1]); phi = rho * np.sum(t[i + 1:])
        dpsi = rho * np.sum(t[: i + 1] / delta[: i + 1]); dphi = rho * np.sum(t[i + 1:] / delta[i + 1:])
        f = 1.0 + psi + phi
        err = 8.0 * EPS * (1.0 + abs(psi) + abs(phi)) * 1.0 + EPS * K * (abs(psi) + abs(phi)) * 0
        if f > 0: hi = min(hi, mu)
        else: lo = max(lo, mu)
        if abs(f) <= EPS * K * (1.0 + abs(psi) + abs(phi)) or hi - lo <= 2 * EPS * max(abs(lo), abs(hi)):
            break
        # BNS rational step: psi ~ a1 + b1/(D_i - mu'), phi ~ a2 + b2/(D_i+1 - mu')
        if i < K - 1:
            Di, Dj = dd[i] - mu, dd[i + 1] - mu   # current distances to the two poles
            b1 = Di * Di * dpsi; a1 = psi - Di * dpsi
            b2 = Dj * Dj * dphi; a2 = phi - Dj * dphi
            c = 1.0 + a1 + a2
            # c + b1/(Di - eta) + b2/(Dj - eta) = 0 -> c (Di-eta)(Dj-eta) + b1 (Dj-eta) + b2 (Di-eta) = 0
            A = c; B = -(c * (Di + Dj) + b1 + b2); C = c * Di * Dj + b1 * Dj + b2 * Di
            disc = B * B - 4 * A * C
            if disc < 0: disc = 0.0
            sq = np.sqrt(disc)
            if A == 0: eta = -C / B if B != 0 else 0.0
            else:
                # want the root with Di < eta < Dj (Di<0<Dj): choose numerically stable form
                q = -0.5 * (B + np.sign(B) * sq) if B != 0 else 0.5 * sq
                r1 = q / A if A != 0 else np.inf
                r2 = C / q if q != 0 else np.inf
                eta = r1 if Di < r1 < Dj else r2
        else:
            # last root: psi ~ a1 + b1/(D_K-1 - eta); phi = 0  (simple one-pole + two-pole fallback)
            Di = dd[K - 1] - mu
            Dm = dd[K - 2] - mu if K > 1 else None
            if K > 1:
                # interpolate psi without last term by pole K-2, last term exact
                tl = rho * z2[K - 1]
                psi0 = psi - tl / Di; dpsi0 = dpsi - tl / (Di * Di)
                b1 = Dm * Dm * dpsi0; a1 = psi0 - Dm * dpsi0
                c = 1.0 + a1
                A = c; B = -(c * (Dm + Di) + b1 + tl); C = c * Dm * Di + b1 * Di + tl * Dm
                disc = max(B * B - 4 * A * C, 0.0); sq = np.sqrt(disc)
                q = -0.5 * (B + np.sign(B) * sq) if B != 0 else 0.5 * sq
                r1 = q / A if A != 0 else np.inf
                r2 = C / q if q != 0 else np.inf
                eta = r1 if r1 > Di else r2   # eta must exceed Di (= -mu + ... <0): lam beyond last pole
            else:
                eta = -f / dpsi
        new = mu + eta
        if not (lo < new < hi) or not np.isfinite(new):
            new = 0.5 * (lo + hi)
        mu = new
    return org, mu


def merge(d1, Q1, d2, Q2, e):
    """Eigen-decomposition of blkdiag(T1', T2') + |e| u u^T given those of T1', T2'."""
    n1, n2 = len(d1), len(d2); n = n1 + n2
    d = np.concatenate([d1, d2])
    z = np.concatenate([Q1[-1, :], np.sign(e) * Q2[0, :] if e != 0 else Q2[0, :]])
    Q = np.zeros((n, n)); Q[:n1, :n1] = Q1; Q[n1:, n1:] = Q2
    z = z / np.sqrt(2.0); rho = 2.0 * abs(e)
    order = np.argsort(d, kind="stable")
    tol = 8.0 * EPS * max(np.max(np.abs(d)), np.max(np.abs(z)))
    keep = []; defl = []
    if rho * np.max(np.abs(z)) <= tol:
        defl = list(order)
    else:
        pj = None
        for j in order:
            if rho * abs(z[j]) <= tol:
                defl.append(j); continue
            if pj is None:
                pj = j; continue
            s = z[pj]; c = z[j]
            tau = np.hypot(c, s); t = d[j] - d[pj]
            c /= tau; s = -s / tau
            if abs(t * c * s) <= tol:
                z[j] = tau; z[pj] = 0.0
                qp = Q[:, pj].copy(); qj = Q[:, j].copy()
                Q[:, pj] = c * qp + s * qj; Q[:, j] = -s * qp + c * qj
                t2 = d[pj] * c * c + d[j] * s * s
                d[j] = d[pj] * s * s + d[j] * c * c
                d[pj] = t2
                defl.append(pj); pj = j
            else:
                keep.append(pj); pj = j
        keep.append(pj)
    K = len(keep)
    lam = np.empty(n); Qn = np.empty((n, n))
    if K:
        dl = d[keep]; w = z[keep]
        # note: after Givens deflation dl may be slightly unsorted? LAPACK keeps order by construction.
        z2 = w * w
        delta = np.empty((K, K)); roots = np.empty(K)
        for i in range(K):
            org, mu = secular_root(i, dl, z2, rho)
            delta[i] = (dl - dl[org]) - mu
            roots[i] = dl[org] + mu
        # Gu-Eisenstat z-hat
        zh = np.empty(K)
        for j in range(K):
            p = -delta[j, j]      # lam_j - d_j
            for i in range(K):
                if i != j:
                    p *= -delta[i, j] / (dl[i] - dl[j])     # (lam_i - d_j)/(d_i - d_j)
            zh[j] = np.sign(w[j]) * np.sqrt(abs(p) / rho) if rho > 0 else w[j]
        U = zh[None, :] / delta   # U[i, j] = zh_j / (d_j - lam_i)
        U /= np.linalg.norm(U, axis=1, keepdims=True)
        Qn[:, :K] = Q[:, keep] @ U.T
        lam[:K] = roots
    Qn[:, K:] = Q[:, defl]; lam[K:] = d[defl]
    o = np.argsort(lam, kind="stable")
    return lam[o], Qn[:, o], K


def stedc(d, e, leaf=32):
    n = len(d)
    if n <= leaf:
        return leaf_eig(d, e)
    k = n // 2
    d = d.copy()
    ek = e[k - 1]
    d[k - 1] -= abs(ek); d[k] -= abs(ek)
    l1, Q1 = stedc(d[:k], e[: k - 1], leaf)
    l2, Q2 = stedc(d[k:], e[k:], leaf)
    lam, Q, K = merge(l1, Q1, l2, Q2, ek)
    return lam, Q


if __name__ == "__main__":
    import scipy.linalg as sl, sys
    rng = np.random.default_rng(0)
    for name in ["random", "graded", "wilkinson", "gram", "clustered", "zeros"]:
        n = 300
        if name == "random":
            d = rng.standard_normal(n); e = rng.standard_normal(n - 1)
        elif name == "graded":
            d = np.logspace(0, -14, n); e = 0.1 * np.sqrt(d[:-1] * d[1:])
        elif name == "wilkinson":
            d = np.abs(np.arange(n) - n // 2).astype(float); e = np.ones(n - 1)
        elif name == "clustered":
            d = np.ones(n) + 1e-13 * rng.standard_normal(n); e = 1e-9 * rng.standard_normal(n - 1)
        elif name == "zeros":
            d = rng.standard_normal(n); e = rng.standard_normal(n - 1); e[::7] = 0; d[10:20] = 1.0; e[10:19] = 0
        else:
            X = rng.standard_normal((2000, 12)); W = rng.standard_normal((12, n))
            P = np.hstack([np.cos(X @ W), np.sin(X @ W)])[:, :n] / np.sqrt(n)
            A = P.T @ P / 2000
            T, Qh = sl.hessenberg(A, calc_q=True)
            d = np.diag(T).copy(); e = np.diag(T, 1).copy()
        lam, Q = stedc(d, e, leaf=16)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        ref = np.linalg.eigvalsh(T)
        nrm = np.max(np.abs(ref))
        print(f"{name:10s} lam err {np.max(np.abs(lam-ref))/nrm:.2e} orth {np.max(np.abs(Q.T@Q-np.eye(n))):.2e} resid {np.max(np.abs(T@Q-Q*lam))/nrm:.2e}")
