"""Primal solve orchestration: the five device stages of `NeoLSSVM._optimize_β̂_γ` behind the C ABI.

Mirrors the reference's `_optimize_β̂_γ` (/root/reference/src/neo_ls_svm/_neo_ls_svm.py:77-189) for
the constant-diagonal complexity matrix all shipped feature maps produce (_feature_maps.py:129-135).
Rows may be sharded over ranks (one process per GPU): the partial Gram / right-hand side and the
per-γ error sums are summed with `torch.distributed.all_reduce` (NCCL over NVLink); the eigensolve and
the m×m Cholesky are replicated.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

N_GAMMAS_PRIMAL = 1024  # _neo_ls_svm.py:146


def gamma_grid(num: int) -> np.ndarray:
    """γ grid, generated with NumPy exactly as the reference does (_neo_ls_svm.py:146, :270)."""
    return np.logspace(np.log10(1e-6), np.log10(20), num, dtype=np.float64)


def _dist_on() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1


def _all_reduce(t: torch.Tensor) -> None:
    if _dist_on():
        torch.distributed.all_reduce(torch.view_as_real(t) if t.is_complex() else t)


def select_gamma(sums: np.ndarray, classifier: bool) -> tuple[int, np.ndarray]:
    """argmin of the selection objective with NumPy semantics (first minimum, first NaN wins).

    sums: 3×G array [Σ s|e|, Σ s(|e|≥1), Σ s·max(0,|e|-1)]; _neo_ls_svm.py:158-165.
    """
    err = sums[0]
    obj = (sums[1] + sums[2]) + err if classifier else err
    return int(np.argmin(obj)), obj


@dataclass
class PrimalFit:
    """Device-resident result of the primal solve for this rank's rows."""

    beta: torch.Tensor  # β̂ (Cholesky re-solve, :178), complex128 (m,)
    beta_eig: torch.Tensor  # β̂ from the eigen-expansion (:175)
    gamma: float
    opt: int
    gammas: np.ndarray
    loo_errors: np.ndarray  # (G,)
    objective: np.ndarray
    A: torch.Tensor  # Hermitian Gram (m, m), summed over ranks
    b: torch.Tensor
    lam: torch.Tensor
    Q: torch.Tensor
    U: torch.Tensor  # upper Cholesky factor of γ*C + A (cho_factor layout)
    inv_c: float
    rows: dict = field(default_factory=dict)  # per-row vectors for the local shard
    loo_score: float = float("nan")
    loo_error: float = float("nan")
    stage_ms: dict = field(default_factory=dict)


def primal_fit(
    X: torch.Tensor,
    y: torch.Tensor,
    s: torch.Tensor,
    shift: torch.Tensor,
    W: torch.Tensor,
    classifier: bool,
    n_global: int | None = None,
    ctx: _lib.Context | None = None,
    time_stages: bool = False,
    stash: bool | str = "auto",
    host_rows: tuple | None = None,
) -> PrimalFit:
    """Run stages 1–4 for the local row shard.

    X (n×d), y (n), s (n, ALREADY divided by the global weight sum, :110), shift (d), W (d×D) are
    contiguous float64 CUDA tensors.  `n_global` is the total number of rows over all ranks.

    `stash`: keep σ²ᵢ(γ_g) for every local row and γ (n×G doubles, 8 KB/row) while sweeping, so that the
    per-row outputs at the selected γ are a column gather instead of a second n·m² projection pass.
    "auto" enables it when the buffer fits in 60% of the free HBM.

    `host_rows=(Xh, yh, sh)`: the rows are still on the host (contiguous float64 CPU tensors, ideally pinned); pass
    X = y = s = None.  They are uploaded underneath the Gram pass (`nls_primal_gram_h2d`) and the device copies
    serve the later passes.
    """
    ctx = ctx or _lib.context((X if X is not None else W).device.index)
    n, d = (X if X is not None else host_rows[0]).shape
    D = W.shape[1]
    m = D + 1
    n_global = int(n_global if n_global is not None else n)
    inv_c = float(n_global) * float(m)  # 1/c with c = 1/φ.size, :117-118
    ev = []

    def mark(name):
        if time_stages:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e))

    mark("start")
    if host_rows is not None:
        A, b, X, y, s = ctx.primal_gram_h2d(*host_rows, shift, W)  # upload + stage 1+2
    else:
        A, b = ctx.primal_gram(X, y, s, shift, W)  # stage 1+2
    _all_reduce(A)
    _all_reduce(b)
    mark("gram")
    lam, Q = ctx.heev(A, inv_c)  # stage 3 (replicated)
    v, _ = ctx.primal_coeffs(Q, lam, b, inv_c)
    mark("eigh")
    gammas_np = gamma_grid(N_GAMMAS_PRIMAL)
    gammas = torch.from_numpy(gammas_np).to(X.device)
    stash_buf = None
    if stash == "auto":
        free_bytes, _ = torch.cuda.mem_get_info(X.device)
        stash = n * N_GAMMAS_PRIMAL * 8 < 0.6 * free_bytes
    if stash:
        stash_buf = torch.empty((n, N_GAMMAS_PRIMAL), dtype=torch.float64, device=X.device)
    sums = ctx.primal_loo_sweep(X, y, s, shift, W, Q, lam, v, inv_c, gammas, classifier, stash=stash_buf)  # stage 4a+4b
    _all_reduce(sums)
    sums_np = sums.cpu().numpy()
    opt, obj = select_gamma(sums_np, classifier)
    gamma = float(gammas_np[opt])
    mark("sweep")
    _, beta_eig = ctx.primal_coeffs(Q, lam, None, inv_c, gamma, v=v)
    U, beta = ctx.cholesky_solve(A, gamma / inv_c, b)  # :177-178
    mark("solve")
    sigma2 = stash_buf[:, opt].contiguous() if stash_buf is not None else None
    del stash_buf
    rows = ctx.primal_finalize(X, y, s, shift, W, Q, lam, inv_c, gamma, beta_eig, beta, classifier, sigma2=sigma2)  # stage 4c
    mark("finalize")
    # LOO score (:171-174) from weighted sums over all ranks.
    yhat_loo = rows["yhat_loo"]
    if classifier:
        t = torch.stack([torch.sum(s * (torch.sign(yhat_loo) == y)), torch.sum(s)])
        _all_reduce(t)
        loo_score = float(t[0] / t[1])
    else:
        t = torch.stack([torch.sum(s * y), torch.sum(s)])
        _all_reduce(t)
        ybar = t[0] / t[1]
        t2 = torch.stack([torch.sum(s * (y - yhat_loo) ** 2), torch.sum(s * (y - ybar) ** 2)])
        _all_reduce(t2)
        loo_score = float(1.0 - t2[0] / t2[1])
    fit = PrimalFit(
        beta=beta, beta_eig=beta_eig, gamma=gamma, opt=opt, gammas=gammas_np, loo_errors=sums_np[0].copy(),
        objective=obj, A=A, b=b, lam=lam, Q=Q, U=U, inv_c=inv_c, rows=rows, loo_score=loo_score,
        loo_error=float(sums_np[0][opt]),
    )
    if time_stages:
        torch.cuda.synchronize()
        fit.stage_ms = {ev[i][0]: ev[i - 1][1].elapsed_time(ev[i][1]) for i in range(1, len(ev))}
    return fit


def variance_weights(lam: torch.Tensor, inv_c: float, gamma: float) -> torch.Tensor:
    """w_k = 1/(c (λ_k + γ)): (γC + A)⁻¹ = Q diag(w) Qᴴ in the eigenbasis of A/c."""
    return inv_c / (lam + gamma)
