"""On-GPU diagnostic: stage-by-stage errors of the CUDA path against the CPU oracle, plus peaks.

Not a pytest module (run with `python tests/gpu_diag.py`); prints one line per check so that a single
gpurun call yields as much evidence as possible.  Writes gpurun_out/diag.json.
"""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from neo_ls_svm_b200 import _lib, _primal  # noqa: E402
from neo_ls_svm_b200.datasets import load_case  # noqa: E402
from oracle import neo_oracle as orc  # noqa: E402

OUT = {}


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = float(np.max(np.abs(b))) or 1.0
    return float(np.max(np.abs(a - b))) / den


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def report(key, val):
    OUT[key] = val
    print(f"[diag] {key} = {val}", flush=True)


def golden(name):
    with np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def stage_checks(ctx, tag, name, chunk=None):
    g = golden(name)
    X, y, sw, Xt, _ = load_case(name)
    classifier = bool(g["classifier"])
    y_ = np.where(y == np.unique(y)[0], -1.0, 1.0) if classifier else y.astype(np.float64)
    s = np.ones(len(y)) if sw is None else sw.astype(np.float64)
    s = s / s.sum()
    shift = g["shift"].ravel()
    W = g["A_map"] / g["scale"].reshape(-1, 1)
    if chunk:
        ctx.set_chunk_rows(chunk)
    Xd, yd, sd, shd, Wd = dev(X), dev(y_), dev(s), dev(shift), dev(W)
    n, d = X.shape
    D = W.shape[1]
    m = D + 1
    # stage 1
    phi = ctx.feature_map(Xd[:300].contiguous(), shd, Wd).cpu().numpy()
    phi_ref = orc.feature_map(X[:300], g["shift"], g["scale"], g["A_map"])
    report(f"{tag}/{name}/phi", rel(phi, phi_ref))
    # stage 2
    A, b = ctx.primal_gram(Xd, yd, sd, shd, Wd)
    phi_all = orc.feature_map(X, g["shift"], g["scale"], g["A_map"])
    Sphi = s[:, None] * phi_all
    A_ref = Sphi.conj().T @ Sphi
    A_ref = (A_ref + A_ref.conj().T) / 2
    b_ref = Sphi.conj().T @ (s * y_)
    report(f"{tag}/{name}/A", rel(A.cpu().numpy(), A_ref))
    report(f"{tag}/{name}/b", rel(b.cpu().numpy(), b_ref))
    # stage 3
    inv_c = float(n) * m
    lam, Q = ctx.heev(A, inv_c)
    lam_ref = np.linalg.eigvalsh(A_ref * inv_c)
    report(f"{tag}/{name}/lam", rel(lam.cpu().numpy(), lam_ref))
    Qn = Q.cpu().numpy()
    report(f"{tag}/{name}/Q_unitary", float(np.max(np.abs(Qn.conj().T @ Qn - np.eye(m)))))
    report(f"{tag}/{name}/Q_resid", rel((A.cpu().numpy() * inv_c) @ Qn, Qn * lam.cpu().numpy()[None, :]))
    # full solve
    fit = _primal.primal_fit(Xd, yd, sd, shd, Wd, classifier, ctx=ctx)
    report(f"{tag}/{name}/opt", [fit.opt, int(g["opt"])])
    report(f"{tag}/{name}/loo_errors", rel(fit.loo_errors, g["loo_errors"]))
    report(f"{tag}/{name}/beta", rel(fit.beta.cpu().numpy(), g["beta"]))
    report(f"{tag}/{name}/beta_eig", rel(fit.beta_eig.cpu().numpy(), g["beta"]))
    for k_mine, k_gold in (("loo_residuals", "loo_residuals"), ("loo_leverage", "loo_leverage"),
                           ("residuals", "residuals"), ("loo_std", "loo_std")):
        report(f"{tag}/{name}/{k_mine}", rel(fit.rows[k_mine].cpu().numpy(), g[k_gold]))
    report(f"{tag}/{name}/loo_score", abs(fit.loo_score - float(g["loo_score"])))
    report(f"{tag}/{name}/L_diag", rel(np.diag(fit.U.cpu().numpy()), g["L_diag"]))
    # predict
    w = _primal.variance_weights(fit.lam, fit.inv_c, fit.gamma)
    yhat, sigma = ctx.primal_predict(dev(Xt), shd, Wd, beta=fit.beta, B=fit.Q, w=w, want_std=True)
    report(f"{tag}/{name}/decision", rel(yhat.cpu().numpy(), g["decision"]))
    report(f"{tag}/{name}/std", rel(sigma.cpu().numpy(), g["std"]))
    q = ctx.quantile_epilogue(
        dev(g["decision"]), dev(g["std"]), dev(g["cqr_abs_beta"]), dev(g["cqr_rel_beta"]), dev(g["cqr_abs_bias"]),
        dev(g["cqr_rel_bias"]), not classifier,
        dev(g["iso_x"]) if classifier else None, dev(g["iso_y"]) if classifier else None)
    report(f"{tag}/{name}/quantiles", rel(q.cpu().numpy(), g["quantiles_accuracy"]))


def peaks(ctx):
    report("dmma_peak_tflops", ctx.dmma_peak_tflops(20000))
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    report("cublas_dgemm_8192_tflops", 2 * n**3 / (best * 1e-3) / 1e12)
    az = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    bz = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    az @ bz
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        az @ bz
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    report("cublas_zgemm_4096_real_tflops", 8 * 4096**3 / (best * 1e-3) / 1e12)
    del a, b, az, bz


def timing(ctx, n=262144, d=64, D=1024):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d))
    y = X[:, :8].sum(axis=1) + rng.standard_normal(n)
    W = rng.standard_normal((d, D)) * 0.3
    s = np.full(n, 1.0 / n)
    Xd, yd, sd, shd, Wd = dev(X), dev(y), dev(s), dev(np.zeros(d)), dev(W)
    ctx.set_chunk_rows(32768)
    _primal.primal_fit(Xd, yd, sd, shd, Wd, False, ctx=ctx)  # warm-up
    ctx.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = _primal.primal_fit(Xd, yd, sd, shd, Wd, False, ctx=ctx, time_stages=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    prof = ctx.profile_read()
    ctx.profile(False)
    report("timing/n", n)
    report("timing/fit_s", t1 - t0)
    report("timing/rows_per_s", n / (t1 - t0))
    report("timing/stage_ms", fit.stage_ms)
    report("timing/kernels", prof)
    m = D + 1
    flops = {"gram": 4.0 * m * m * n, "project": 8.0 * m * m * n, "sweep": 4.0 * m * 1024 * n, "variance": 8.0 * m * m * n}
    report("timing/kernel_tflops", {k: flops[k] / (prof[k]["ms"] * 1e-3) / 1e12 for k in flops if prof[k]["ms"] > 0})


def eig_timing(ctx):
    """Hand-written block Jacobi vs cuSOLVER Zheevd on Gram matrices of the hot path."""
    rng = np.random.default_rng(0)
    for m in (513, 1025, 2049):
        n = 4 * m
        Z = rng.standard_normal((n, m - 1)) * 1.5
        phi = np.concatenate([np.exp(-1j * Z) / np.sqrt(m - 1), np.ones((n, 1))], axis=1) / n
        A = phi.conj().T @ phi
        A = (A + A.conj().T) / 2
        Ad = dev(A)
        scale = float(n) * m
        ref = np.linalg.eigvalsh(A * scale)
        for kind in ("cusolver", "jacobi"):
            ctx.set_eigensolver(kind)
            ctx.heev(Ad, scale)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lam, Q = ctx.heev(Ad, scale)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            Qn = Q.cpu().numpy()
            report(f"eig/{kind}/m{m}", {
                "ms": dt * 1e3, "sweeps": ctx.last_eig_sweeps(),
                "lam_err": float(np.max(np.abs(lam.cpu().numpy() - ref)) / ref[-1]),
                "resid": float(np.max(np.abs((A * scale) @ Qn - Qn * lam.cpu().numpy()[None, :])) / ref[-1]),
                "unitary": float(np.max(np.abs(Qn.conj().T @ Qn - np.eye(m))))})
    ctx.set_eigensolver("jacobi")


def eig_small(ctx):
    """Padding / odd-size correctness of the Jacobi solver (graded and clustered spectra included)."""
    rng = np.random.default_rng(1)
    ctx.set_eigensolver("jacobi")
    for m in (1, 2, 3, 15, 16, 17, 33, 100, 257, 640):
        B = rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m))
        Qr, _ = np.linalg.qr(B)
        lam0 = np.logspace(0, -18, m) if m % 2 else np.repeat(rng.standard_normal((m + 3) // 4), 4)[:m]
        A = (Qr * lam0) @ Qr.conj().T
        A = (A + A.conj().T) / 2
        ref = np.linalg.eigvalsh(A)
        lam, Q = ctx.heev(dev(A), 1.0)
        lam, Qn = lam.cpu().numpy(), Q.cpu().numpy()
        nrm = np.abs(ref).max()
        report(f"eig_small/m{m}", {
            "sweeps": ctx.last_eig_sweeps(), "lam_err": float(np.max(np.abs(lam - ref)) / nrm),
            "resid": float(np.max(np.abs(A @ Qn - Qn * lam[None, :])) / nrm),
            "unitary": float(np.max(np.abs(Qn.conj().T @ Qn - np.eye(m))))})


def main():
    which = sys.argv[1:] or ["peaks", "tma", "plain", "timing"]
    print("device:", torch.cuda.get_device_name(0), flush=True)
    if "plain" in which:
        os.environ["NLS_NO_TMA"] = "1"
        ctx_plain = _lib.Context(0)
        os.environ["NLS_NO_TMA"] = "0"
        for name in ("reg_small", "clf_small"):
            try:
                stage_checks(ctx_plain, "plain", name, chunk=512)
            except Exception as exc:  # noqa: BLE001
                report(f"plain/{name}/EXC", repr(exc))
    ctx = _lib.Context(0)
    if "peaks" in which:
        peaks(ctx)
    if "tma" in which:
        for name, chunk in (("reg_small", 512), ("clf_small", 640), ("c1", 4096), ("c2_small", 32768), ("c3_small", 2048)):
            try:
                stage_checks(ctx, "tma", name, chunk=chunk)
            except Exception as exc:  # noqa: BLE001
                report(f"tma/{name}/EXC", repr(exc))
    if "eig" in which:
        try:
            eig_timing(ctx)
        except Exception as exc:  # noqa: BLE001
            report("eig/EXC", repr(exc))
    if "eig_small" in which:
        try:
            eig_small(ctx)
        except Exception as exc:  # noqa: BLE001
            report("eig_small/EXC", repr(exc))
    if "timing" in which:
        try:
            timing(ctx)
        except Exception as exc:  # noqa: BLE001
            report("timing/EXC", repr(exc))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(OUT, f, indent=1, default=str)


if __name__ == "__main__":
    main()
