"""NumPy prototype of the blocked Hermitian tridiagonalisation as the cooperative panel kernel computes it
(design study for csrc/hetrd.cuh): per column ONE exchange of the raw updated column x and ONE reduction of
[x^H x, x^H y_raw, x^H a1, V^H x, W^H x] — everything that depends on the Householder scalars (beta, tau) is derived
from those sums afterwards (deferred-alpha form)."""
import numpy as np


def hetrd_blocked(A, nb=32):
    A = A.copy(); n = A.shape[0]
    cplx = np.iscomplexobj(A)
    d = np.zeros(n); e = np.zeros(n - 1); tau = np.zeros(n - 1, dtype=A.dtype)
    Vall = np.zeros((n, n), dtype=A.dtype)   # column c: reflector c (unit at row c+1)
    Ts = []
    for k0 in range(0, n - 1, nb):
        jb = min(nb, n - 1 - k0)
        V = np.zeros((n, jb), dtype=A.dtype); W = np.zeros((n, jb), dtype=A.dtype)
        T = np.zeros((jb, jb), dtype=A.dtype)
        for j in range(jb):
            c = k0 + j
            # pass 1 (per row slab): raw updated column
            a = A[c:, c] - V[c:, :j] @ W[c, :j].conj() - W[c:, :j] @ V[c, :j].conj()
            d[c] = a[0].real
            x = a[1:]                              # rows c+1..n-1  -> exchanged (barrier B1)
            # pass 2 (per row slab): hemv with the un-updated trailing matrix + partial sums
            At = A[c + 1:, c + 1:]
            y_raw = At @ x
            a1 = At[:, 0]                          # first column of the trailing matrix
            s_xx = np.vdot(x, x).real; s_xy = np.vdot(x, y_raw); s_xa = np.vdot(x, a1)
            Vx = V[c + 1:, :j].conj().T @ x; Wx = W[c + 1:, :j].conj().T @ x
            a11 = At[0, 0].real
            v0 = V[c + 1, :j]; w0 = W[c + 1, :j]
            # ---- barrier B2: everything below is computed redundantly by every CTA from the sums ----
            alpha = x[0]
            xnorm2 = s_xx - abs(alpha) ** 2
            if xnorm2 <= 0 and (not cplx or alpha.imag == 0):
                e[c] = alpha.real; tau[c] = 0
                v = np.zeros_like(x); v[0] = 1
                V[c + 1:, j] = v; W[c + 1:, j] = 0; Vall[c + 1:, c] = v
                continue
            beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + max(xnorm2, 0.0)), alpha.real)
            t = (beta - alpha) / beta
            sigma = 1.0 / (alpha - beta)
            e[c] = beta; tau[c] = t
            Vv = sigma * (Vx - beta * v0.conj())       # V^H v,  v = sigma (x - beta e1)
            Wv = sigma * (Wx - beta * w0.conj())       # W^H v
            # v^H A v (real) = |sigma|^2 [x^H y_raw - 2 beta Re(x^H a1) + beta^2 a11]
            vAv = abs(sigma) ** 2 * (s_xy.real - 2 * beta * s_xa.real + beta * beta * a11)
            # y = A v - V (W^H v) - W (V^H v);   y^H v = v^H A v - 2 Re((W^H v)^H (V^H v))
            yHv = vAv - 2 * np.vdot(Wv, Vv).real
            alpha2 = -0.5 * t * (np.conj(t) * yHv)     # w = tau y + alpha2 v,  alpha2 = -1/2 tau ((tau y)^H v)
            v = sigma * x; v[0] = 1.0
            y = sigma * (y_raw - beta * a1) - V[c + 1:, :j] @ Wv - W[c + 1:, :j] @ Vv
            w = t * y + alpha2 * v
            V[c + 1:, j] = v; W[c + 1:, j] = w; Vall[c + 1:, c] = v
            T[:j, j] = -t * (T[:j, :j] @ Vv); T[j, j] = t   # compact WY factor (CTA 0)
        r = k0 + jb
        A[r:, r:] -= V[r:, :] @ W[r:, :].conj().T + W[r:, :] @ V[r:, :].conj().T   # her2k, full storage
        Ts.append(T)
    d[n - 1] = A[n - 1, n - 1].real
    return d, e, tau, Vall, Ts


def backtransform(Z, Vall, Ts, nb=32):
    """X = Q Z with Q = prod_p (I - V_p T_p V_p^H), applied panel by panel from the last, on Y = X^T."""
    n = Z.shape[0]
    Y = Z.T.astype(Vall.dtype).copy()
    starts = list(range(0, n - 1, nb))
    for p in reversed(range(len(starts))):
        k0 = starts[p]; jb = Ts[p].shape[0]
        Vp = Vall[:, k0:k0 + jb]
        C1 = Y @ Vp.conj()              # GEMM1  (n x jb)
        C2 = C1 @ Ts[p].T               # small
        Y -= C2 @ Vp.T                  # GEMM2
    return Y.T


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for cplx in (True, False):
        for n in (2, 5, 33, 100, 257):
            M = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0)
            A = M @ M.conj().T / n
            A = (A + A.conj().T) / 2
            d, e, tau, Vall, Ts = hetrd_blocked(A, nb=32)
            T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
            lam, Z = np.linalg.eigh(T)
            X = backtransform(Z, Vall, Ts)
            ref = np.linalg.eigvalsh(A)
            print(cplx, n, "lam", np.max(np.abs(lam - ref)) / ref[-1], "resid", np.max(np.abs(A @ X - X * lam)) / ref[-1],
                  "orth", np.max(np.abs(X.conj().T @ X - np.eye(n))))
