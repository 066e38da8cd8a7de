"""Accuracy and timing of the stage-3 eigensolvers on the GPU box, side by side (-> gpurun_out/eig_bench.json):

    python scripts/eig_bench.py [--complex 513,1025,2049,4097] [--real 4096,16384] [--kinds dc,jacobi,cusolver]

`dc`      hand-written tridiagonalisation + divide and conquer + back-transformation (csrc/hetrd.cuh, csrc/stedc.cuh)
`jacobi`  hand-written two-sided block Jacobi (csrc/jacobi_wide.cuh)
`cusolver` library comparator (Zheevd / Dsyevd)
Complex Hermitian matrices go through `nls_heev`; real symmetric ones through the dual fit's eigen-stage, timed as
`nls_dual_sweep` on a kernel-matrix-like input (the only entry point that solves a real symmetric problem).
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from neo_ls_svm_b200 import _lib  # noqa: E402


def gram_like(m, rng, complex_=True):
    """Hermitian PSD matrix with the fast-decaying spectrum of a random-feature Gram matrix (14 decades)."""
    U, _ = np.linalg.qr(rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if complex_ else 0))
    lam = np.logspace(0, -14, m)
    A = (U * lam) @ U.conj().T
    return (A + A.conj().T) / 2


def time_call(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--complex", default="513,1025,2049,4097")
    ap.add_argument("--real", default="4096,16384")
    ap.add_argument("--kinds", default="dc,jacobi,cusolver")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "eig_bench.json"))
    args = ap.parse_args()
    kinds = args.kinds.split(",")
    rng = np.random.default_rng(0)
    out = {"complex": {}, "real": {}}
    code = {"jacobi": 0, "cusolver": 1, "auto": 2, "dc": 3}
    for m in [int(x) for x in args.complex.split(",") if x]:
        A = gram_like(m, rng)
        Ad = torch.from_numpy(A).cuda()
        ref = np.linalg.eigvalsh(A)
        res = {}
        for kind in kinds:
            if kind == "jacobi" and m > 2100:
                continue
            ctx = _lib.Context(0)
            _lib.check(ctx.lib.nls_ctx_set_eigensolver(ctx.handle, code[kind]))
            try:
                ms = time_call(lambda: ctx.heev(Ad, 1.0))
                lam, Q = ctx.heev(Ad, 1.0)
            except Exception as exc:  # noqa: BLE001
                res[kind] = {"error": str(exc)}
                print(m, kind, "ERROR", exc, flush=True)
                continue
            lam_h, Qh = lam.cpu().numpy(), Q.cpu().numpy()
            r = {
                "ms": ms,
                "lam_err": float(np.max(np.abs(lam_h - ref)) / ref[-1]),
                "orth": float(np.max(np.abs(Qh.conj().T @ Qh - np.eye(m)))),
                "resid": float(np.max(np.abs(A @ Qh - Qh * lam_h[None, :])) / ref[-1]),
            }
            res[kind] = r
            print(f"complex m={m:5d} {kind:9s} {ms:9.2f} ms  lam {r['lam_err']:.1e} orth {r['orth']:.1e} resid {r['resid']:.1e}", flush=True)
            del ctx
        out["complex"][m] = res
    for n in [int(x) for x in args.real.split(",") if x]:
        p = 48
        Xt = rng.standard_normal((n, p)) * 0.35
        y = rng.standard_normal(n)
        s = np.full(n, 1.0 / n)
        sn = s / np.median(s)
        gam = np.logspace(-6, np.log10(20), 128)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        Xd, yd, sd, snd, gd = up(Xt), up(y), up(s), up(sn), up(gam)
        res = {}
        lam_ref = None
        for kind in kinds:
            if kind == "jacobi":
                continue
            ctx = _lib.Context(0)
            _lib.check(ctx.lib.nls_ctx_set_eigensolver(ctx.handle, code[kind]))
            try:
                ms = time_call(lambda: ctx.dual_sweep(Xd, yd, sd, snd, gd, False), reps=2)
                sums, yl, lam = ctx.dual_sweep(Xd, yd, sd, snd, gd, False)
            except Exception as exc:  # noqa: BLE001
                res[kind] = {"error": str(exc)}
                print(n, kind, "ERROR", exc, flush=True)
                continue
            lam_h = lam.cpu().numpy()
            if lam_ref is None and n <= 4096:
                d2 = np.sum(Xt * Xt, 1)[:, None] - 2 * Xt @ Xt.T + np.sum(Xt * Xt, 1)[None, :]
                F = np.exp(-0.5 * np.maximum(d2, 0)) + 1.0
                lam_ref = np.linalg.eigvalsh(sn[:, None] * F * sn[None, :])
            r = {"dual_sweep_ms": ms, "sums0": float(sums[0, 64])}
            if lam_ref is not None:
                r["lam_err"] = float(np.max(np.abs(lam_h - lam_ref)) / lam_ref[-1])
            res[kind] = r
            print(f"real    n={n:5d} {kind:9s} dual_sweep {ms:9.1f} ms  {r}", flush=True)
            del ctx
            torch.cuda.empty_cache()
        if "dc" in res and "cusolver" in res and "sums0" in res["dc"] and "sums0" in res["cusolver"]:
            res["sums_rel_diff"] = abs(res["dc"]["sums0"] - res["cusolver"]["sums0"]) / abs(res["cusolver"]["sums0"])
        out["real"][n] = res
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
