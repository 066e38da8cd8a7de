"""C5 of BASELINE.json at full size, under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tests/gpu_c5_full.py [--rows 16000000]

Large-feature primal fit n = 16M, d = 128, num_features = 4096 (m = 4097) with rows sharded over the ranks, then
batched predict / predict_std / quantile epilogue over every local row.  No CPU oracle can hold this shape
(BASELINE.md §2), so correctness is checked through size-independent properties of the result:

  * A is Hermitian, trace(A/c) = Σλ, and β̂ solves the normal equations (γ*C + A) β̂ = b;
  * every rank selects the same γ index; leverages lie in [0, 1);
  * LOO identity per row: loo_residuals · (1 − leverage) = residuals (the two come from different kernels);
  * fit/predict consistency per row: loo_std² = σ² + (s σ²)²/(1 − leverage) with σ from the predict_std kernel
    (triangular U⁻¹ basis) while loo_std comes from the fit's eigenbasis sweep; ŷ from predict = y + residuals;
  * quantile epilogue: monotone in the quantile for monotone coefficients.

Not a pytest module (needs 8 GPUs); writes gpurun_out/c5_full.json on rank 0.
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neo_ls_svm_b200 import OrthogonalRandomFourierFeatures, _lib, _primal  # noqa: E402
from neo_ls_svm_b200.datasets import fast_regression_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16_000_000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--features", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=262_144)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, D = args.rows, args.d, args.features
    m = D + 1
    r0, r1 = rank * n // world, (rank + 1) * n // world
    out = {"world": world, "n": n, "d": d, "m": m, "rows_per_gpu": r1 - r0}

    def say(k, v):
        out[k] = v
        if rank == 0:
            print(f"[c5] {k} = {v}", flush=True)

    def gmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # Host pre-pass (identical on every rank): supervised affine map + ORF frequencies from the first 100k rows.
    Xp, yp = fast_regression_rows(n, d, d // 2, row_begin=0, row_end=min(100_000, n))
    fm = OrthogonalRandomFourierFeatures(num_features=D).fit(Xp, yp, np.ones(len(yp)))
    shift, W = fm.device_weights(d)
    del Xp, yp
    t0 = time.perf_counter()
    X, y = fast_regression_rows(n, d, d // 2, row_begin=r0, row_end=r1)
    say("datagen_s", time.perf_counter() - t0)
    s = np.full(r1 - r0, 1.0 / n)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
    Xd, yd, sd, shd, Wd = up(X), up(y), up(s), up(shift), up(W)
    del X
    ctx = _lib.context(local)
    # warm-up on a slice (scratch allocation, graph instantiation, NCCL channels)
    w_rows = min(65_536, r1 - r0)
    _primal.primal_fit(Xd[:w_rows], yd[:w_rows], sd[:w_rows] * (n / (w_rows * world)), shd, Wd, False,
                       n_global=w_rows * world, ctx=ctx)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fit = _primal.primal_fit(Xd, yd, sd, shd, Wd, False, n_global=n, ctx=ctx, time_stages=True)
    e1.record()
    torch.cuda.synchronize()
    fit_ms = gmax(e0.elapsed_time(e1))
    say("fit_ms", fit_ms)
    say("fit_rows_per_s", n / (fit_ms * 1e-3))
    flops_row = 2.0 * d * D + 12.0 * m * m + 4.0 * m * 1024
    say("fit_tflops_total", n * flops_row / (fit_ms * 1e-3) / 1e12)
    say("stage_ms", {k: round(v, 2) for k, v in fit.stage_ms.items()})
    say("eig_sweeps", ctx.last_eig_sweeps())

    # ---- invariants -----------------------------------------------------------------------------
    opts = torch.tensor([fit.opt], dtype=torch.int64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(opts) for _ in range(world)]
        dist.all_gather(gathered, opts)
        opts = torch.cat(gathered)
    say("gamma_index", int(fit.opt))
    say("gamma_index_equal_on_all_ranks", bool((opts == opts[0]).all().item()))
    A, b, beta = fit.A, fit.b, fit.beta
    say("A_hermitian_rel", float((A - A.conj().T).abs().max() / A.abs().max()))
    say("trace_identity_rel", float(abs(torch.diagonal(A).real.sum() * fit.inv_c - fit.lam.sum()) / fit.lam.sum()))
    lhs = A @ beta + (fit.gamma / fit.inv_c) * beta
    say("normal_equations_rel", float((lhs - b).abs().max() / b.abs().max()))
    rows = fit.rows
    lev = rows["loo_leverage"]
    say("leverage_min_max", [gmax(-float(lev.min())) * -1.0, gmax(float(lev.max()))])
    loo_id = (rows["loo_residuals"] * (1.0 - lev) - rows["residuals"]).abs().max() / rows["residuals"].abs().max()
    say("loo_identity_rel", gmax(float(loo_id)))
    say("loo_score", fit.loo_score)

    # ---- batched predict / predict_std / quantiles over every local row -------------------------------
    U = torch.triu(fit.U)
    B = torch.linalg.solve_triangular(U, torch.eye(m, dtype=torch.complex128, device=dev), upper=True).contiguous()
    w = torch.ones(m, dtype=torch.float64, device=dev)
    Q3 = np.array([0.025, 0.5, 0.975])
    # Monotone synthetic coefficients of the two coherent quantile regressors ([sigma, |yhat|, 1] x Q); the real ones
    # come from two host LPs on 1440 calibration rows (out of the device hot path).
    beta_abs = up(np.array([[-1.96, 0.0, 1.96], [0.0, 0.0, 0.0], [-0.1, 0.0, 0.1]]))
    beta_rel = up(np.array([[-0.01, 0.0, 0.01], [0.0, 0.0, 0.0], [-0.02, 0.0, 0.02]]))
    bias = up(np.zeros(3))
    yhat_all = torch.empty(r1 - r0, dtype=torch.float64, device=dev)
    sig_all = torch.empty(r1 - r0, dtype=torch.float64, device=dev)
    mono_ok = True
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for i0 in range(0, r1 - r0, args.batch):
        i1 = min(r1 - r0, i0 + args.batch)
        yh, sg = ctx.primal_predict(Xd[i0:i1], shd, Wd, beta=beta, B=B, w=w, want_std=True, b_upper=True)
        yhat_all[i0:i1], sig_all[i0:i1] = yh, sg
        q = ctx.quantile_epilogue(yh, sg, beta_abs, beta_rel, bias, bias, True)
        mono_ok = mono_ok and bool((q[:, 1:] >= q[:, :-1]).all().item())
    e1.record()
    torch.cuda.synchronize()
    pred_ms = gmax(e0.elapsed_time(e1))
    say("predict_std_quantiles_ms", pred_ms)
    say("predict_rows_per_s", n / (pred_ms * 1e-3))
    say("predict_tflops_total", n * (2.0 * d * D + 4.0 * D + 4.0 * m * m) / (pred_ms * 1e-3) / 1e12)
    say("quantiles_monotone", bool(gmax(0.0 if mono_ok else 1.0) == 0.0))
    yh_id = (yhat_all - (yd + rows["residuals"])).abs().max() / yd.abs().max()
    say("predict_equals_fit_yhat_rel", gmax(float(yh_id)))
    s2 = sig_all**2
    loo_var = s2 + (sd * s2) ** 2 / (1.0 - lev)
    std_id = (torch.sqrt(loo_var) - rows["loo_std"]).abs().max() / rows["loo_std"].abs().max()
    say("predict_std_equals_fit_loo_std_rel", gmax(float(std_id)))
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "c5_full.json"), "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
