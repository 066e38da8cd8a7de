"""On-GPU runs of the other BASELINE.json configurations (C2, C4, C5-shaped) with timings and parity.

    python tests/gpu_configs.py [c2] [c4] [c5]        -> gpurun_out/configs.json

These are parity-test cases, not bench lines (bench.py is C3).  Reference CPU numbers for context are in
BASELINE.md §2 (C2: fit 31.8 s, predict_proba 0.77 s, predict_interval 7.8 s first / 2.1 s cached).
"""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from neo_ls_svm_b200 import NeoLSSVM, _lib, _primal  # noqa: E402
from neo_ls_svm_b200.datasets import make_churn_rows, make_regression_rows  # noqa: E402
from oracle import neo_oracle as orc  # noqa: E402

OUT = {}


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) or 1.0))


def report(k, v):
    OUT[k] = v
    print(f"[cfg] {k} = {v}", flush=True)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def c2():
    """C2: churn-shaped binary classification n=100k d=70 (default ORF D=512), predict_proba + predict_interval."""
    X, y = make_churn_rows(115_000, 70, 20)
    Xtr, ytr, Xte = X[:100_000], y[:100_000], X[100_000:]
    NeoLSSVM().fit(Xtr[:3000], ytr[:3000])  # warm-up (library load, numba JIT of the host pre-pass)
    model, t_fit = timed(lambda: NeoLSSVM().fit(Xtr, ytr))
    report("c2/fit_s", t_fit)
    report("c2/fit_rows_per_s", len(ytr) / t_fit)
    # the device part alone (stages 1-4c given the fitted map)
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device)
    shift, W = model.primal_feature_map_.device_weights(70)
    y_ = np.where(ytr == model.classes_[0], -1.0, 1.0)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    args = (up(Xtr), up(y_), up(np.full(len(y_), 1.0 / len(y_))), up(shift), up(W))
    _primal.primal_fit(*args, True, ctx=ctx)
    fit, t_dev = timed(lambda: _primal.primal_fit(*args, True, ctx=ctx))
    report("c2/device_solve_s", t_dev)
    report("c2/gamma_index", fit.opt)
    _, t = timed(lambda: model.predict_proba(Xte))
    report("c2/predict_proba_15k_s", t)
    _, t = timed(lambda: model.predict_std(Xte))
    report("c2/predict_std_15k_s", t)
    iv, t1 = timed(lambda: model.predict_interval(Xte, coverage=0.95))
    _, t2 = timed(lambda: model.predict_interval(Xte, coverage=0.95))
    report("c2/predict_interval_15k_first_s", t1)
    report("c2/predict_interval_15k_cached_s", t2)
    report("c2/interval_shape", list(iv.shape))
    # parity vs the chunked CPU oracle on the same fitted map
    aff = model.primal_feature_map_.affine_feature_map
    ref, t_cpu = timed(lambda: orc.primal_fit_chunked(Xtr, y_, np.ones(len(y_)), aff.shift_, aff.scale_, aff.A_, True, chunk=8192))
    report("c2/oracle_chunked_cpu_s", t_cpu)
    report("c2/opt_equal", [int(np.argmin(np.abs(model.γs_ - model.γ_))), ref["opt"]])
    report("c2/beta_rel", rel(model.β̂_, ref["beta"]))
    report("c2/loo_errors_rel", rel(model.loo_errors_γs_, ref["loo_errors"]))
    report("c2/loo_residuals_rel", rel(model.loo_residuals_, ref["loo_residuals"]))
    report("c2/loo_std_rel", rel(model.loo_std_, ref["loo_std"]))
    phi = orc.feature_map(Xte[:2000], aff.shift_, aff.scale_, aff.A_)
    report("c2/decision_rel", rel(model.decision_function(Xte[:2000]), orc.primal_decision(phi, ref["beta"])))
    report("c2/std_rel", rel(model.predict_std(Xte[:2000]), orc.primal_std(phi, ref["L"])))


def c4():
    """C4: dual-space solve; full size n=16384 d=32 timed, parity vs the einsum-free oracle at n=3000."""
    X, y = make_regression_rows(3400, 32, n_informative=16)
    m, t = timed(lambda: NeoLSSVM(dual=True).fit(X[:3000], y[:3000]))
    report("c4/n3000_fit_s", t)
    ref = orc.dual_fit(m.X_, y[:3000].astype(np.float64), np.ones(3000), False)
    report("c4/n3000_opt", [int(np.argmin(np.abs(m.γs_ - m.γ_))), ref["opt"]])
    report("c4/n3000_alpha_rel", rel(m.α̂_, ref["alpha"]))
    report("c4/n3000_loo_errors_rel", rel(m.loo_errors_γs_, ref["loo_errors"]))
    report("c4/n3000_loo_residuals_rel", rel(m.loo_residuals_, ref["loo_residuals"]))
    report("c4/n3000_loo_std_rel", rel(m.loo_std_, ref["loo_std"]))
    Xq = orc.affine_map(X[3000:], m.dual_feature_map_.shift_, m.dual_feature_map_.scale_, m.dual_feature_map_.A_)
    report("c4/n3000_decision_rel", rel(m.decision_function(X[3000:]), orc.dual_decision(Xq, m.X_, ref["alpha"])))
    report("c4/n3000_std_rel", rel(m.predict_std(X[3000:]), orc.dual_std(Xq, m.X_, ref["L"])))
    del m
    X, y = make_regression_rows(16_384 + 2000, 32, n_informative=16)
    model, t = timed(lambda: NeoLSSVM(dual=True).fit(X[:16_384], y[:16_384]))
    report("c4/n16384_fit_s", t)
    report("c4/n16384_gamma_index", int(np.argmin(np.abs(model.γs_ - model.γ_))))
    report("c4/n16384_loo_score", float(model.loo_score_))
    # oracle-free invariant: (γ S⁻² + F) α̂ = y  <=>  residuals_ = F α̂ − y = −γ α̂ / sn²
    s = np.full(16_384, 1.0 / 16_384)
    sn = s / np.median(np.abs(s))
    report("c4/n16384_normal_eq_rel", rel(model.residuals_, -float(model.γ_) * model.α̂_ / sn**2))
    _, t = timed(lambda: model.predict_std(X[16_384:]))
    report("c4/n16384_predict_std_2k_s", t)
    _, t = timed(lambda: model.predict(X[16_384:]))
    report("c4/n16384_predict_2k_s", t)


def c5():
    """C5-shaped: d=128, num_features=4096 (m=4097) at n=24k; parity vs the chunked CPU oracle + timing."""
    n, d, D = 24_000, 128, 4096
    X, y = make_regression_rows(n, d, n_informative=64)
    rng = np.random.default_rng(0)
    shift = np.zeros(d)
    scale = np.ones((1, d))
    A_map = rng.standard_normal((d, D)) * 0.12
    ctx = _lib.context()
    dev = torch.device("cuda", ctx.device)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    args = (up(X), up(y), up(np.full(n, 1.0 / n)), up(shift), up(A_map))
    _primal.primal_fit(*args, False, ctx=ctx)
    fit, t = timed(lambda: _primal.primal_fit(*args, False, ctx=ctx, time_stages=True))
    report("c5/n24k_device_solve_s", t)
    report("c5/stage_ms", fit.stage_ms)
    ref, t_cpu = timed(lambda: orc.primal_fit_chunked(X, y, np.ones(n), shift, scale, A_map, False, chunk=4000))
    report("c5/oracle_cpu_s", t_cpu)
    report("c5/opt", [fit.opt, ref["opt"]])
    report("c5/A_rel", rel(fit.A.cpu().numpy(), ref["A"]))
    report("c5/lam_rel", rel(fit.lam.cpu().numpy(), ref["lam"]))
    report("c5/beta_rel", rel(fit.beta.cpu().numpy(), ref["beta"]))
    report("c5/loo_errors_rel", rel(fit.loo_errors, ref["loo_errors"]))
    report("c5/loo_residuals_rel", rel(fit.rows["loo_residuals"].cpu().numpy(), ref["loo_residuals"]))
    report("c5/loo_std_rel", rel(fit.rows["loo_std"].cpu().numpy(), ref["loo_std"]))
    w = _primal.variance_weights(fit.lam, fit.inv_c, fit.gamma)
    (yh, sg), t = timed(lambda: ctx.primal_predict(args[0], args[3], args[4], beta=fit.beta, B=fit.Q, w=w, want_std=True))
    report("c5/predict_std_24k_s", t)
    phi = orc.feature_map(X[:1500], shift, scale, A_map)
    report("c5/decision_rel", rel(yh[:1500].cpu().numpy(), orc.primal_decision(phi, ref["beta"])))
    report("c5/std_rel", rel(sg[:1500].cpu().numpy(), orc.primal_std(phi, ref["L"])))


def main():
    which = sys.argv[1:] or ["c2", "c4", "c5"]
    for name in which:
        try:
            {"c2": c2, "c4": c4, "c5": c5}[name]()
        except Exception as exc:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            report(f"{name}/EXC", repr(exc))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(OUT, f, indent=1, default=str)


if __name__ == "__main__":
    main()
