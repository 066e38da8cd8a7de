"""Design study (CPU, NumPy): can the γ sweep go onto the INT8 core?  The sweep contracts P (or U) with
rγ[k, g] = 1/(γ_g + λ_k), whose dynamic range along k is anti-correlated with P's, so a plain fixed-point split fails.
Grouped two-sided scaling: for a group of γ with reference γ_ref = the group's smallest γ,
    P'_ik = P_ik / (λ_k + γ_ref),   r'_kg = (λ_k + γ_ref) / (λ_k + γ_g)  in (γ_ref/γ_g, 1],
then both operands are quantised to 56 fractional bits relative to their row maxima (what csrc/ozaki.cuh does; the
integer products are exact, so quantising the operands and multiplying in FP64 emulates the scheme to FP64 rounding).

    python scripts/sweep_int8_study.py [case ...]
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy.linalg import eigh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neo_ls_svm_b200.datasets import load_case  # noqa: E402
from oracle import neo_oracle as orc  # noqa: E402

BITS = 56


def quantise_rows(X):
    amax = np.max(np.abs(X), axis=1)
    m, ex = np.frexp(np.where(amax > 0, amax, 1.0))
    e = np.where(m <= 0.99, ex + 1, ex + 2)
    scale = np.ldexp(1.0, BITS - e)
    return np.rint(X * scale[:, None]) / scale[:, None]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def study(cases, group_counts=(0, 1, 8), verbose=True):
    """results[name][groups] = {"opt_equal", "curve", "num", "den", "loo", "elementwise"} against the all-FP64 sweep."""
    results = {}
    for name in cases:
        if name.startswith("ragged"):  # the shapes of tests/test_gpu_primal.py::test_ragged_shapes_match_oracle
            _, n_, d_, D_ = name.split(":")
            n_, d_, D_ = int(n_), int(d_), int(D_)
            rng = np.random.default_rng(n_ + d_ + D_)
            X = rng.standard_normal((n_, d_))
            y = np.sin(X[:, 0]) + 0.3 * rng.standard_normal(n_)
            sw = rng.uniform(0.5, 1.5, n_)
            sw[::17] = 0.0
            g = {"shift": rng.standard_normal(d_) * 0.1, "scale": rng.uniform(0.5, 2.0, (1, d_)), "A_map": rng.standard_normal((d_, D_)) * 0.7,
                 "classifier": False, "opt": -1}
        else:
            with np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")) as z:
                g = {k: z[k] for k in z.files}
            X, y, sw, _, _ = load_case(name)
        classifier = bool(g["classifier"])
        y_ = np.where(y == np.unique(y)[0], -1.0, 1.0) if classifier else y.astype(np.float64)
        s = np.ones(len(y)) if sw is None else sw.astype(np.float64)
        s = s / s.sum()
        n = len(y)
        phi = orc.feature_map(X, g["shift"], g["scale"], g["A_map"])
        m = phi.shape[1]
        Sphi = s[:, None] * phi
        A = Sphi.conj().T @ Sphi
        A = (A + A.conj().T) / 2
        b = Sphi.conj().T @ (s * y_)
        c = 1.0 / (n * m)
        lam, Q = eigh(A / c)
        v = (Q.conj().T @ b) / c
        T = phi @ Q
        P = np.real(T * v[None, :])
        U = (T.real**2 + T.imag**2) / c
        gammas = orc.gamma_grid(orc.N_GAMMAS_PRIMAL)
        rg = 1 / (gammas[None, :] + lam[:, None])
        num0, den0 = P @ rg, U @ rg

        def curve(num, den):
            with np.errstate(divide="ignore", invalid="ignore"):
                loo = (num - y_[:, None]) / (1 - (s[:, None] ** 2) * den)
            if classifier:
                orc._clip_classifier(loo, y_)
            a = np.abs(loo)
            err = s @ a
            obj = err + (s @ (a >= 1)) + (s @ np.maximum(0, a - 1)) if classifier else err
            return loo, err, int(np.argmin(obj))

        loo0, err0, opt0 = curve(num0, den0)
        if verbose:
            print(f"== {name}: n={n} m={m} classifier={classifier} opt={opt0} (golden {int(g['opt'])}); λ in [{lam.min():.2e}, {lam.max():.2e}]")
        results[name] = {}
        for groups in group_counts:
            num, den = np.empty_like(num0), np.empty_like(den0)
            if groups == 0:  # plain fixed point: no column scaling at all
                rq = quantise_rows(rg.T).T
                num, den = quantise_rows(P) @ rq, quantise_rows(U) @ rq
            for gi in range(groups):
                sl = slice(gi * len(gammas) // groups, (gi + 1) * len(gammas) // groups)
                gref = gammas[sl][0]
                d = np.abs(lam) + gref  # λ can be slightly negative at rounding level: keep the scaling positive
                Pq, Uq = quantise_rows(P / d[None, :]), quantise_rows(U / d[None, :])
                rq = quantise_rows((d[:, None] / (gammas[None, sl] + lam[:, None])).T).T
                num[:, sl], den[:, sl] = Pq @ rq, Uq @ rq
            loo, err, opt = curve(num, den)
            results[name][groups] = {
                "opt_equal": opt == opt0, "curve": rel(err, err0), "num": rel(num, num0), "den": rel(den, den0),
                "loo": rel(loo[:, opt0], loo0[:, opt0]),
                "elementwise": float(np.max(np.abs(loo[:, opt0] - loo0[:, opt0]) / (1e-9 * np.abs(loo0[:, opt0]) + 1e-12 * np.max(np.abs(loo0[:, opt0]))))),
            }
            if verbose:
                print(f"   groups={groups:2d}: opt={opt}  LOO curve rel {rel(err, err0):.1e}  num rel {rel(num, num0):.1e}  den rel {rel(den, den0):.1e}  "
                      f"LOO residuals at opt rel {rel(loo[:, opt0], loo0[:, opt0]):.1e}  elementwise worst |Δ|/(1e-9|ref|+1e-12 max) "
                      f"{float(np.max(np.abs(loo[:, opt0] - loo0[:, opt0]) / (1e-9 * np.abs(loo0[:, opt0]) + 1e-12 * np.max(np.abs(loo0[:, opt0]))))):.2f}")
    return results


def main():
    study(sys.argv[1:] or ["c1", "c3_small", "clf_small", "reg_small", "ragged:130:3:40", "ragged:1000:7:100", "ragged:2049:5:129"])


if __name__ == "__main__":
    main()
