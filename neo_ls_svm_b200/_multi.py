"""Rows sharded over several GPUs of one box, behind the unchanged `NeoLSSVM(...).fit / predict*` API.

The reference is single-process NumPy (/root/reference/src/neo_ls_svm/_neo_ls_svm.py:379-404); SURVEY.md §8e shards the
primal solve by rows: every GPU computes the partial Gram / right-hand side and the per-γ error sums of its rows, the
partials are summed with NCCL all-reduce over NVLink, the small m×m eigensolve runs once and its result is broadcast.

Device selection keeps the estimator's signature unchanged:

    NLS_DEVICES=0,1,2,3   (or "all")          environment, read at call time
    neo_ls_svm_b200.set_devices([0, 1, 2, 3])  module-level setter (None: back to the environment / single GPU)

One host thread drives all GPUs: every stage call only enqueues work on that GPU's stream, so the shards run
concurrently.  (The multi-process form — one rank per GPU under torchrun with `torch.distributed` all-reduces — is what
`bench.py --gpus N` measures; both go through the same C entry points.)  The dual path is "replicas only".
"""

from __future__ import annotations

import os

import numpy as np

from . import _lib
from ._primal import N_GAMMAS_PRIMAL, PrimalFit, gamma_grid, select_gamma

_DEVICES: list[int] | None = None


def set_devices(devices) -> None:
    """GPUs the primal solve and the batched predict paths shard their rows over (None: NLS_DEVICES / one GPU)."""
    global _DEVICES
    _DEVICES = None if devices is None else [int(d) for d in devices]


def devices() -> list[int]:
    """The selected GPUs; the first one hosts the replicated small solves and the supervised pre-pass."""
    import torch

    if _DEVICES is not None:
        devs = list(_DEVICES)
    else:
        env = os.environ.get("NLS_DEVICES", "").strip()
        if not env:
            return [torch.cuda.current_device()] if torch.cuda.is_available() else [0]
        devs = list(range(torch.cuda.device_count())) if env.lower() == "all" else [int(t) for t in env.split(",") if t.strip()]
    count = torch.cuda.device_count()
    bad = [d for d in devs if d < 0 or d >= count]
    if bad or not devs or len(set(devs)) != len(devs):
        raise ValueError(f"NLS_DEVICES / set_devices: invalid device list {devs} ({count} visible)")
    return devs


def shard_bounds(n: int, parts: int) -> list[tuple[int, int]]:
    """Contiguous row ranges, the same split bench.py uses for its ranks."""
    return [(k * n // parts, (k + 1) * n // parts) for k in range(parts)]


def _all_reduce_sum(tensors) -> None:
    """In-place sum over one tensor per device: NCCL (single-process group call) over NVLink."""
    import torch
    import torch.cuda.nccl as nccl

    views = [torch.view_as_real(t) if t.is_complex() else t for t in tensors]
    nccl.all_reduce(views)


def _broadcast(t, devs):
    import torch

    return [t if t.device.index == d else t.to(torch.device("cuda", d), non_blocking=True) for d in devs]


def primal_fit_sharded(X, y, s_norm, shift, W, classifier: bool, devs: list[int], X_primary=None) -> PrimalFit:
    """Stages 1–4c with the rows split over `devs`.  X (n×d), y, s_norm (already divided by the total weight) are host
    arrays; X_primary (optional) is a copy of X that already lives on devs[0].  Returns a PrimalFit whose per-row
    vectors are HOST tensors (`rows["_stacked"]`: 5×n) and whose small tensors live on devs[0]."""
    import torch

    n, d = X.shape
    D = W.shape[1]
    m = D + 1
    inv_c = float(n) * float(m)
    N = len(devs)
    bounds = shard_bounds(n, N)
    ctxs = [_lib.context(dv) for dv in devs]
    tdev = [torch.device("cuda", dv) for dv in devs]
    y64 = np.ascontiguousarray(y, dtype=np.float64)
    Xs, ys, ss, sh, Ws = [], [], [], [], []
    for k, (r0, r1) in enumerate(bounds):
        if X_primary is not None:
            Xk = X_primary[r0:r1] if k == 0 else X_primary[r0:r1].to(tdev[k], non_blocking=True)
        else:
            Xk = torch.from_numpy(np.ascontiguousarray(X[r0:r1], dtype=np.float64)).to(tdev[k], non_blocking=True)
        Xs.append(Xk.contiguous())
        ys.append(torch.from_numpy(y64[r0:r1]).to(tdev[k], non_blocking=True))
        ss.append(torch.from_numpy(np.ascontiguousarray(s_norm[r0:r1])).to(tdev[k], non_blocking=True))
        sh.append(torch.from_numpy(shift).to(tdev[k], non_blocking=True))
        Ws.append(torch.from_numpy(W).to(tdev[k], non_blocking=True))
    # stage 1+2 on every GPU, then the first exchange step
    AB = [ctxs[k].primal_gram(Xs[k], ys[k], ss[k], sh[k], Ws[k]) for k in range(N)]
    _all_reduce_sum([ab[0] for ab in AB])
    _all_reduce_sum([ab[1] for ab in AB])
    A, b = AB[0]
    # stage 3 once, broadcast
    lam, Q = ctxs[0].heev(A, inv_c)
    v, _ = ctxs[0].primal_coeffs(Q, lam, b, inv_c)
    lams, Qs, vs = _broadcast(lam, devs), _broadcast(Q, devs), _broadcast(v, devs)
    gammas_np = gamma_grid(N_GAMMAS_PRIMAL)
    sums, stashes = [], []
    for k in range(N):
        gam = torch.from_numpy(gammas_np).to(tdev[k])
        free_bytes, _ = torch.cuda.mem_get_info(tdev[k])
        nk = Xs[k].shape[0]
        stash = torch.empty((nk, N_GAMMAS_PRIMAL), dtype=torch.float64, device=tdev[k]) if nk * N_GAMMAS_PRIMAL * 8 < 0.6 * free_bytes else None
        stashes.append(stash)
        sums.append(ctxs[k].primal_loo_sweep(Xs[k], ys[k], ss[k], sh[k], Ws[k], Qs[k], lams[k], vs[k], inv_c, gam, classifier, stash=stash))
    _all_reduce_sum(sums)  # second exchange step: 3×G error sums
    sums_np = sums[0].cpu().numpy()
    opt, obj = select_gamma(sums_np, classifier)
    gamma = float(gammas_np[opt])
    _, beta_eig = ctxs[0].primal_coeffs(Q, lam, None, inv_c, gamma, v=v)
    U, beta = ctxs[0].cholesky_solve(A, gamma / inv_c, b)
    betas_eig, betas = _broadcast(beta_eig, devs), _broadcast(beta, devs)
    rows_host = torch.empty((5, n), dtype=torch.float64).pin_memory()
    score_terms = []
    for k, (r0, r1) in enumerate(bounds):
        sigma2 = stashes[k][:, opt].contiguous() if stashes[k] is not None else None
        stashes[k] = None
        rows = ctxs[k].primal_finalize(Xs[k], ys[k], ss[k], sh[k], Ws[k], Qs[k], lams[k], inv_c, gamma, betas_eig[k], betas[k],
                                       classifier, sigma2=sigma2)
        rows_host[:, r0:r1].copy_(rows["_stacked"], non_blocking=True)
        yl, sk, yk = rows["yhat_loo"], ss[k], ys[k]
        if classifier:
            score_terms.append(torch.stack([torch.sum(sk * (torch.sign(yl) == yk)), torch.sum(sk)]))
        else:
            score_terms.append((sk, yk, yl))
    for dv in devs:
        torch.cuda.synchronize(dv)
    if classifier:
        t = sum(st.cpu() for st in score_terms)
        loo_score = float(t[0] / t[1])
    else:
        sw = sum(float(torch.sum(sk)) for sk, _, _ in score_terms)
        ybar = sum(float(torch.sum(sk * yk)) for sk, yk, _ in score_terms) / sw
        num = sum(float(torch.sum(sk * (yk - yl) ** 2)) for sk, yk, yl in score_terms)
        den = sum(float(torch.sum(sk * (yk - ybar) ** 2)) for sk, yk, _ in score_terms)
        loo_score = 1.0 - num / den
    torch.cuda.set_device(devs[0])
    return PrimalFit(
        beta=beta, beta_eig=beta_eig, gamma=gamma, opt=opt, gammas=gammas_np, loo_errors=sums_np[0].copy(), objective=obj,
        A=A, b=b, lam=lam, Q=Q, U=U, inv_c=inv_c, rows={"_stacked": rows_host}, loo_score=loo_score,
        loo_error=float(sums_np[0][opt]),
    )


def primal_predict_sharded(X, shift, W, beta, B, w, devs: list[int], want_decision: bool, want_std: bool):
    """decision_function / predict_std with the query rows split over `devs` (no collective: rows are independent).
    shift, W, beta, B, w are tensors on devs[0]; returns host NumPy vectors."""
    import torch

    n = X.shape[0]
    bounds = shard_bounds(n, len(devs))
    outs = []
    for k, (r0, r1) in enumerate(bounds):
        if r1 == r0:
            outs.append((None, None))
            continue
        dv = torch.device("cuda", devs[k])
        ctx = _lib.context(devs[k])
        mv = lambda t: None if t is None else t.to(dv, non_blocking=True)  # noqa: E731
        Xk = torch.from_numpy(np.ascontiguousarray(X[r0:r1], dtype=np.float64)).to(dv, non_blocking=True)
        outs.append(ctx.primal_predict(Xk, mv(shift), mv(W), beta=mv(beta) if want_decision else None,
                                       B=mv(B) if want_std else None, w=mv(w) if want_std else None, want_std=want_std, b_upper=True))
    yhat = np.concatenate([o[0].cpu().numpy() for o in outs if o[0] is not None]) if want_decision else None
    sigma = np.concatenate([o[1].cpu().numpy() for o in outs if o[1] is not None]) if want_std else None
    torch.cuda.set_device(devs[0])
    return yhat, sigma
