"""Design study (CPU only, NumPy): how many 7-bit slices would an INT8 (tcgen05 kind::i8) emulation of the three
FP64 GEMM stages need to keep the north star's parity (same γ index; β̂, LOO curve, LOO residuals ≤ 1e-9)?

    python scripts/ozaki_study.py [case ...]        (golden cases of tests/golden, default: c1 c3_small clf_small)

Ozaki scheme: every operand vector along K (a row of A, a row of Bᵀ) is scaled by a power of two to |x| < 1 and cut
into s signed 7-bit digits, x = Σ_p q_p 128^-(p+1).  A digit-plane product A_p·B_qᵀ is an exact small-integer GEMM
(INT8×INT8→INT32 on the tensor cores; emulated here in float64, which is exact for these magnitudes), and
C = Σ_{p+q<s} 128^-(p+q+2) A_p·B_qᵀ is recombined in FP64 — s(s+1)/2 integer GEMMs per FP64 GEMM.  The script replaces
the Gram, projection and sweep products of the reference algorithm (oracle/neo_oracle.py, chunked form) by this
emulation and reports the parity against the all-FP64 oracle per number of slices.

This is NOT product code and not a checker: it only informs DESIGN.md §9 (1).
"""

from __future__ import annotations

import os
import sys

import numpy as np
from scipy.linalg import cho_factor, cho_solve, eigh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neo_ls_svm_b200.datasets import load_case  # noqa: E402
from oracle import neo_oracle as orc  # noqa: E402


RADIX = float(os.environ.get("OZAKI_RADIX", "128"))  # 128: signed 7-bit digits; 256: 8-bit digits (first signed, rest unsigned on hardware)


def slices(X: np.ndarray, s: int):
    """Digit planes of the rows of X: X ≈ scale[:, None] * Σ_p planes[p] * RADIX^-(p+1), |digits| ≤ RADIX / 2."""
    amax = np.max(np.abs(X), axis=1)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)  # |x| / 2^e ≤ 1/2
    scale = np.exp2(e)
    r = X / scale[:, None]
    planes = []
    for _ in range(s):
        r = r * RADIX
        q = np.rint(r)
        planes.append(q)
        r = r - q
    return planes, scale


def ozaki_mm(A: np.ndarray, B: np.ndarray, s: int) -> np.ndarray:
    """A (M×K) @ B (N×K)ᵀ through s digit planes per operand (s = 0: plain FP64)."""
    if s == 0:
        return A @ B.T
    Ap, sa = slices(A, s)
    Bp, sb = slices(B, s)
    C = np.zeros((A.shape[0], B.shape[0]))
    for t in range(s - 1, -1, -1):  # p + q = t, smallest terms first
        acc = np.zeros_like(C)
        for p in range(t + 1):
            acc += Ap[p] @ Bp[t - p].T  # exact: integer-valued float64, |sum| < 2^53
        C += acc * RADIX ** -(t + 2)
    return C * sa[:, None] * sb[None, :]


def fit(X, y, s_w, shift, scale, A_map, classifier, n_slices):
    """The reference algorithm in its chunked single-T form with the three GEMM stages emulated."""
    n, m = X.shape[0], A_map.shape[1] + 1
    D = m - 1
    s_w = s_w / np.sum(s_w)
    phi = orc.feature_map(X, shift, scale, A_map)
    Cw, Sw = s_w[:, None] * phi[:, :D].real, -s_w[:, None] * phi[:, :D].imag  # weighted cos / sin planes
    mm = lambda a, b: ozaki_mm(a, b, n_slices)  # noqa: E731
    # Stage 2: Re A = CᵀC + SᵀS, Im A = SᵀC − CᵀS (operands K-contiguous: feature rows of the transposed chunk)
    CT, ST = np.ascontiguousarray(Cw.T), np.ascontiguousarray(Sw.T)
    ReA = mm(CT, CT) + mm(ST, ST)
    SC = mm(ST, CT)
    ImA = SC - SC.T
    A = np.zeros((m, m), dtype=np.complex128)
    A[:D, :D] = ReA + 1j * ImA
    A[:D, D] = (s_w @ Cw) + 1j * (s_w @ Sw)  # border: FP64 column sums, as on the GPU
    A[D, :D] = np.conj(A[:D, D])
    A[D, D] = np.sum(s_w**2)
    A = (A + A.conj().T) / 2
    b = np.concatenate([(s_w * y) @ Cw + 1j * ((s_w * y) @ Sw), [np.sum(s_w**2 * y)]])
    c = 1.0 / (n * m)
    lam, Q = eigh(A / c)
    v = (Q.conj().T @ b) / c
    # Stage 4a: T = φQ through the planar [c|s] form
    Cp, Sp = phi[:, :D].real, -phi[:, :D].imag
    Qr, Qi = np.ascontiguousarray(Q[:D].real.T), np.ascontiguousarray(Q[:D].imag.T)
    Tr = mm(Cp, Qr) + mm(Sp, Qi) + Q[D].real[None, :]
    Ti = mm(Cp, Qi) - mm(Sp, Qr) + Q[D].imag[None, :]
    P = Tr * v.real[None, :] - Ti * v.imag[None, :]
    U = (Tr**2 + Ti**2) / c
    # Stage 4b: the two sweep GEMMs
    gammas = orc.gamma_grid(orc.N_GAMMAS_PRIMAL)
    rgT = np.ascontiguousarray((1 / (gammas[None, :] + lam[:, None])).T)
    num, den = mm(P, rgT), mm(U, rgT)
    with np.errstate(divide="ignore", invalid="ignore"):
        loo = (num - y[:, None]) / (1 - (s_w[:, None] ** 2) * den)
    if classifier:
        orc._clip_classifier(loo, y)
    a = np.abs(loo)
    err = s_w @ a
    obj = err + (s_w @ (a >= 1)) + (s_w @ np.maximum(0, a - 1)) if classifier else err
    opt = int(np.argmin(obj))
    beta = cho_solve(cho_factor(gammas[opt] * c * np.eye(m) + A), b)
    return {"opt": opt, "loo_errors": err, "beta": beta, "loo_residuals": loo[:, opt], "A": A, "lam": lam}


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def main():
    cases = sys.argv[1:] or ["c1", "c3_small", "clf_small"]
    print(f"digit radix {int(RADIX)}")
    for name in cases:
        with np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")) as z:
            g = {k: z[k] for k in z.files}
        X, y, sw, _, _ = load_case(name)
        classifier = bool(g["classifier"])
        y_ = np.where(y == np.unique(y)[0], -1.0, 1.0) if classifier else y.astype(np.float64)
        s_w = np.ones(len(y)) if sw is None else sw.astype(np.float64)
        args = (X, y_, s_w, g["shift"], g["scale"], g["A_map"], classifier)
        ref = fit(*args, 0)
        margin = np.sort(ref["loo_errors"] if not classifier else ref["loo_errors"])[:2]
        print(f"== {name}: n={len(y)} m={g['A_map'].shape[1] + 1} classifier={classifier}  reference γ index {ref['opt']} "
              f"(golden {int(g['opt'])}), FP64 vs golden β̂ {rel(ref['beta'], g['beta']):.1e}, "
              f"best-vs-second LOO error gap {(margin[1] - margin[0]) / margin[0]:.1e}")
        print("   slices  int-GEMMs  γ index   A rel      λ rel      β̂ rel     LOO-curve rel  LOO-resid rel")
        for s in ((4, 5, 6, 7, 8, 9) if RADIX == 128 else (4, 5, 6, 7)):
            out = fit(*args, s)
            print(f"   {s:6d}  {s * (s + 1) // 2:9d}  {out['opt']:7d}   {rel(out['A'], ref['A']):.1e}    {rel(out['lam'], ref['lam']):.1e}    "
                  f"{rel(out['beta'], ref['beta']):.1e}    {rel(out['loo_errors'], ref['loo_errors']):.1e}        "
                  f"{rel(out['loo_residuals'], ref['loo_residuals']):.1e}", flush=True)


if __name__ == "__main__":
    main()
