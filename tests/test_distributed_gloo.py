"""N>1 orchestration on CPU (gloo, world_size 2): row sharding, the two all-reduces and γ selection.

The CUDA stages need a GPU, so here each rank produces its shard's partial Gram / per-γ error sums with
the CPU oracle and pushes them through the same `_primal._all_reduce` / `select_gamma` code the GPU path
uses; the result must equal the single-process solve."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(7)
    n, d, D = 600, 4, 24
    X = rng.standard_normal((n, d))
    y = np.sin(X[:, 0]) + 0.3 * rng.standard_normal(n)
    s = rng.uniform(0.5, 1.5, n)
    shift, scale = np.zeros(d), np.ones((1, d))
    A_map = rng.standard_normal((d, D))
    return X, y, s / s.sum(), shift, scale, A_map


def _worker(rank, world, port, out):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scipy.linalg import eigh

    from neo_ls_svm_b200 import _primal
    from neo_ls_svm_b200.datasets import fast_regression_rows
    from oracle import neo_oracle as orc

    X, y, s, shift, scale, A_map = _problem()
    n, m = len(y), A_map.shape[1] + 1
    r0, r1 = rank * n // world, (rank + 1) * n // world
    phi = orc.feature_map(X[r0:r1], shift, scale, A_map)
    Sphi = s[r0:r1, None] * phi
    A = torch.from_numpy(Sphi.conj().T @ Sphi)
    b = torch.from_numpy(Sphi.conj().T @ (s[r0:r1] * y[r0:r1]))
    _primal._all_reduce(A)  # all-reduce #1: partial Gram and right-hand side
    _primal._all_reduce(b)
    A_np = A.numpy()
    A_np = (A_np + A_np.conj().T) / 2
    c = 1.0 / (n * m)
    lam, Q = eigh(A_np / c)
    v = (Q.conj().T @ b.numpy()) / c
    gammas = _primal.gamma_grid(_primal.N_GAMMAS_PRIMAL)
    rg = 1 / (gammas[None, :] + lam[:, None])
    T = phi @ Q
    loo = (np.real(T * v[None, :]) @ rg - y[r0:r1, None]) / (1 - ((s[r0:r1, None] ** 2) * np.abs(T) ** 2 / c) @ rg)
    sums = torch.zeros((3, len(gammas)), dtype=torch.float64)
    sums[0] = torch.from_numpy(s[r0:r1] @ np.abs(loo))
    _primal._all_reduce(sums)  # all-reduce #2: per-γ error sums
    opt, _ = _primal.select_gamma(sums.numpy(), classifier=False)
    # every rank can materialise exactly its own rows of the global benchmark dataset
    Xa, ya = fast_regression_rows(1000, 3, 2, row_begin=0, row_end=1000)
    Xr, yr = fast_regression_rows(1000, 3, 2, row_begin=rank * 500, row_end=(rank + 1) * 500)
    same_rows = np.array_equal(Xa[rank * 500 : (rank + 1) * 500], Xr) and np.array_equal(ya[rank * 500 : (rank + 1) * 500], yr)
    if rank == 0:
        out["A"], out["b"], out["opt"], out["err"] = A_np, b.numpy(), opt, sums[0].numpy()
    out[f"rows{rank}"] = same_rows
    dist.destroy_process_group()


def test_two_rank_sharded_solve_equals_single_process():
    from oracle import neo_oracle as orc

    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        out = dict(out)
    X, y, s, shift, scale, A_map = _problem()
    ref = orc.primal_fit_chunked(X, y, s, shift, scale, A_map, classifier=False, chunk=128)
    assert np.max(np.abs(out["A"] - ref["A"])) < 1e-15
    assert np.max(np.abs(out["b"] - ref["b"])) < 1e-15
    assert out["opt"] == ref["opt"]
    assert np.max(np.abs(out["err"] - ref["loo_errors"])) < 1e-12 * np.max(ref["loo_errors"])
    assert out["rows0"] and out["rows1"]


def test_select_gamma_numpy_semantics():
    from neo_ls_svm_b200._primal import select_gamma

    sums = np.zeros((3, 6))
    sums[0] = [3.0, 1.0, 1.0, 2.0, 5.0, 4.0]
    assert select_gamma(sums, False)[0] == 1  # first minimum wins ties
    sums[0, 4] = np.nan
    assert select_gamma(sums, False)[0] == 4  # np.argmin: the first NaN wins
    sums[0, 4] = 5.0
    sums[1] = [0, 2, 0, 0, 0, 0]
    sums[2] = [0, 0, 0.5, 0, 0, 0]
    assert select_gamma(sums, True)[0] == 2  # classifier objective adds the count and hinge terms
