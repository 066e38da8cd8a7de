"""Benchmark of the Neo LS-SVM primal fit hot path (BASELINE.json: "NeoLSSVM fit rows/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config C3 of BASELINE.json): regression, n = 4,000,000 rows, d = 64, num_features = 1024
(m = 1025), full 1024-γ leave-one-out sweep, uniform sample weights.  The n rows are a fixed global
dataset sharded contiguously over the N ranks ("strong" scaling); the partial Gram / right-hand side and
the per-γ error sums are all-reduced with NCCL, the eigensolve is replicated.

One "step" = one full primal solve (stages 1–4c: feature map → Hermitian Gram → eigendecomposition →
LOO γ sweep → per-row LOO outputs) given the fitted affine/Fourier map (W, shift).  The supervised
affine pre-pass that produces W is host code outside the hot path and is fitted once
on the first 100k rows before timing.

* `value`  : rows/s with X, y, s resident in HBM when the timed region starts.
* `e2e`    : rows/s with HOST (pinned) X, y, s copied H2D and all results copied D2H inside every step; the upload
  goes through `nls_primal_gram_h2d`, which streams row groups on a copy stream underneath the Gram pass.
* `roofline`: tensor roofline of the dominant GEMM stage (and, under `stages`, of all three), timed with CUDA events
  around each launch: FP64 DMMA stages against the DMMA register-loop peak, INT8 (Ozaki) stages against the
  tcgen05 kind::i8 resident-tile peak, both measured in the same run.
* `cpu_baseline`: the CPU oracle port of the reference algorithm on a bounded row sample, host cores.
* `--impl reference`: the reference's CPU algorithm (oracle port, all host threads) on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NeoLSSVM primal fit rows/s"
N_ROWS, N_FEATURES, N_INFORMATIVE, NUM_RFF, N_GAMMAS = 4_000_000, 64, 32, 1024, 1024
PREPASS_ROWS = 100_000
CPU_SAMPLE_ROWS = 100_000  # the reference CPU path is timed on the first 100,000 rows


def flops_per_row(d: int, D: int, G: int) -> float:
    """Algorithmic FP64 flops per training row: 2dD + 12m² + 4mG."""
    m = D + 1
    return 2.0 * d * D + 12.0 * m * m + 4.0 * m * G


def fit_feature_map(d: int, D: int, rows: int):
    """Host pre-pass on the first `rows` rows of the global dataset -> (shift, W)."""
    from neo_ls_svm_b200 import OrthogonalRandomFourierFeatures
    from neo_ls_svm_b200.datasets import fast_regression_rows

    X, y = fast_regression_rows(N_ROWS, d, N_INFORMATIVE, row_begin=0, row_end=rows)
    fm = OrthogonalRandomFourierFeatures(num_features=D).fit(X, y, np.ones(len(y)))
    return fm.device_weights(d)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # Keep only samples under load (clock above idle) for the median.
        busy = [c for c in sm if c > 500] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _load_datasets_module():
    """neo_ls_svm_b200/datasets.py loaded as a stand-alone module (NumPy only): the reference arm must not import
    the product package, let alone its CUDA library."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_nls_bench_datasets", os.path.join(ROOT, "neo_ls_svm_b200", "datasets.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bench_map(d: int, D: int):
    """(shift, W) shared by both arms: the REFERENCE's OrthogonalRandomFourierFeatures(1024) fitted on the first
    100,000 rows of the bench dataset (tests/golden/bench_c3_map.npz, written by `oracle/gen_golden.py bench_map`)."""
    with np.load(os.path.join(ROOT, "tests", "golden", "bench_c3_map.npz")) as z:
        shift, W = z["shift"], z["W"]
    assert shift.shape == (d,) and W.shape == (d, D), "the committed map is for the C3 shape"
    return np.ascontiguousarray(shift), np.ascontiguousarray(W)


def _import_reference():
    """The unmodified reference package, if this machine has it (`baseline/_ref` or /root/reference/src)."""
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nls_numba_cache")
    sys.dont_write_bytecode = True
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(cand, "neo_ls_svm")):
            sys.path.insert(0, cand)
            try:
                import neo_ls_svm  # noqa: F401

                return cand
            except Exception:  # noqa: BLE001
                sys.path.remove(cand)
    return None


def make_cpu_step(rows: int):
    """One pass of the reference's CPU hot path (feature-map transform + `_optimize_β̂_γ`) over the first `rows` rows
    of the C3 dataset, given the fitted map.  Returns (step, kind, detail): the unmodified reference when it is
    importable on this machine, else the oracle port (the reference is pure Python and cannot travel)."""
    ds = _load_datasets_module()
    X, y = ds.fast_regression_rows(N_ROWS, N_FEATURES, N_INFORMATIVE, row_begin=0, row_end=rows)
    s = np.ones(rows)
    shift, W = bench_map(N_FEATURES, NUM_RFF)
    where = _import_reference()
    if where is not None:
        from neo_ls_svm import NeoLSSVM
        from neo_ls_svm._feature_maps import OrthogonalRandomFourierFeatures

        fm = OrthogonalRandomFourierFeatures(num_features=NUM_RFF).fit(X[:2048], y[:2048], s[:2048])  # untimed; map overwritten
        aff = fm.affine_feature_map
        aff.shift_, aff.scale_, aff.A_ = shift.reshape(1, -1).copy(), np.ones((1, N_FEATURES)), W.copy()
        model = NeoLSSVM(primal_feature_map=fm, dual=False)
        model._estimator_type = "regressor"

        def step():
            phi = fm.transform(X)  # _neo_ls_svm.py:386
            C = fm.complexity_matrix.astype(phi.dtype)  # :401
            return model._optimize_β̂_γ(φ=phi, y=y, s=s, C=C)  # :402

        return step, "reference", f"unmodified neo_ls_svm from {where}"
    from oracle import neo_oracle as orc

    def step():
        phi = orc.fourier_map((X - shift[None, :]) @ W)  # _affine_feature_map.py:88 + _feature_maps.py:197-203
        return orc.primal_fit_materialised(phi, y, s, classifier=False)

    return step, "port", "oracle/neo_oracle.py (NumPy/SciPy restatement of the reference, operation for operation)"


def cpu_reference_rows_per_s(rows: int, steps: int, warmup: int):
    """Time `steps` passes after `warmup` untimed ones -> (rows/s, BLAS threads, kind, detail)."""
    from threadpoolctl import threadpool_info

    step, kind, detail = make_cpu_step(rows)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    return rows * steps / dt, threads, kind, detail


def _all_host_threads() -> None:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is a CPU measurement and gets every host
    core at every N, so the interpreter is re-executed once with the thread caps lifted (before NumPy loads)."""
    want = str(os.cpu_count() or 1)
    if os.environ.get("NLS_REF_THREADS_SET") == "1":
        return
    env = dict(os.environ)
    env["NLS_REF_THREADS_SET"] = "1"
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        env[var] = want
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _all_host_threads()
    rows = args.cpu_rows
    warmup = min(args.warmup, 1)
    t0 = time.perf_counter()
    value, threads, kind, detail = cpu_reference_rows_per_s(rows, args.steps, warmup)
    ms_per_step = 1e3 * rows / value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 primal fit n=4M d={N_FEATURES} m={NUM_RFF} G={N_GAMMAS} (reference CPU algorithm, "
                               f"bounded sample of {rows} rows per step as BASELINE.md §3 prescribes; cost is linear in n)",
                   "feature_map": "tests/golden/bench_c3_map.npz (the map both arms use)"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": kind,
                         "sample": f"first {rows} rows of the C3 dataset, feature-map transform + _optimize_β̂_γ; {detail}",
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
        "native_imported": sorted(m for m in sys.modules if m.startswith("neo_ls_svm_b200")),
    }
    print(json.dumps(line))


def other_configs(ctx, peak_tflops: float, hbm_gbs: float, i8_peak: float = 0.0) -> dict:
    """The other BASELINE.json configurations and stage 5 on one GPU, each with the CPU time of the reference algorithm
    (oracle port on a bounded sample) beside it.  Parity-test cases in tests/; these are their timings."""
    import torch

    from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures, _primal
    from neo_ls_svm_b200.datasets import fast_regression_rows, make_churn_rows, make_regression_rows
    from oracle import neo_oracle as orc

    dev = torch.device("cuda", ctx.device)
    out: dict = {}
    int8 = os.environ.get("NLS_GEMM", "ozaki") != "dmma"

    def timed(fn, reps=1):
        torch.cuda.synchronize()
        best = float("inf")
        res = None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return res, best

    def cpu_timed(fn):
        t0 = time.perf_counter()
        res = fn()
        return res, time.perf_counter() - t0

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

    # ---- C1: regression n = 10k, d = 20, default ORF (D = 512), public fit + predict ----
    X, y = make_regression_rows(12_000, 20, n_informative=10)
    NeoLSSVM().fit(X[:2000], y[:2000])  # warm-up (numba JIT of the host pre-pass)
    _, t_fit_first = timed(lambda: NeoLSSVM().fit(X[:10_000], y[:10_000]))  # first call at this shape: allocates device scratch
    m1, t_fit = timed(lambda: NeoLSSVM().fit(X[:10_000], y[:10_000]))
    _, t_pred = timed(lambda: m1.predict(X[10_000:]))
    aff = m1.primal_feature_map_.affine_feature_map
    _, t_cpu = cpu_timed(lambda: orc.primal_fit_materialised(
        orc.feature_map(X[:10_000], aff.shift_, aff.scale_, aff.A_), y[:10_000], np.ones(10_000), False))
    out["c1"] = {"fit_s": t_fit, "fit_first_call_s": t_fit_first, "predict_2k_s": t_pred, "gamma_index": int(np.argmin(np.abs(m1.γs_ - m1.γ_))),
                 "cpu_reference_solve_s": t_cpu, "cpu_sample": "all 10,000 rows, transform + _optimize_β̂_γ (oracle port)"}
    # ---- C2: churn-shaped classifier n = 100k, d = 70, predict_proba + predict_interval on 15k rows ----
    X, y = make_churn_rows(115_000, 70, 20)
    Xtr, ytr, Xte = X[:100_000], y[:100_000], X[100_000:]
    _, t_fit_first = timed(lambda: NeoLSSVM().fit(Xtr, ytr))  # first call at this shape: allocates device scratch
    m2, t_fit = timed(lambda: NeoLSSVM().fit(Xtr, ytr))
    _, t_proba = timed(lambda: m2.predict_proba(Xte))
    _, t_std_first = timed(lambda: m2.predict_std(Xte))  # builds U^-1 from L_ once
    _, t_std = timed(lambda: m2.predict_std(Xte))
    _, t_int1 = timed(lambda: m2.predict_interval(Xte, coverage=0.95))
    _, t_int2 = timed(lambda: m2.predict_interval(Xte, coverage=0.95))
    aff = m2.primal_feature_map_.affine_feature_map
    y_ = np.where(ytr == m2.classes_[0], -1.0, 1.0)
    rows_cpu = 20_000
    _, t_cpu = cpu_timed(lambda: orc.primal_fit_materialised(
        orc.feature_map(Xtr[:rows_cpu], aff.shift_, aff.scale_, aff.A_), y_[:rows_cpu], np.ones(rows_cpu), True))
    out["c2"] = {"fit_s": t_fit, "fit_first_call_s": t_fit_first, "fit_rows_per_s": 100_000 / t_fit, "predict_proba_15k_s": t_proba, "predict_std_15k_s": t_std,
                 "predict_std_15k_first_s": t_std_first,
                 "predict_interval_15k_first_s": t_int1, "predict_interval_15k_cached_s": t_int2,
                 "cpu_reference_solve_rows_per_s": rows_cpu / t_cpu,
                 "cpu_sample": f"first {rows_cpu} rows, transform + _optimize_β̂_γ (oracle port); the reference's own fit of "
                               "all 100,000 rows takes 31.8 s on 8 cores (BASELINE.md)"}
    del m2
    # ---- C4: dual solve n = 16,384, d = 32 ----
    X, y = make_regression_rows(16_384 + 2000, 32, n_informative=16)
    NeoLSSVM(dual=True).fit(X[:1500], y[:1500])
    m4, t_fit_first = timed(lambda: NeoLSSVM(dual=True).fit(X[:16_384], y[:16_384]))  # allocates ~45 GB of scratch
    m4, t_fit = timed(lambda: NeoLSSVM(dual=True).fit(X[:16_384], y[:16_384]))
    _, t_std = timed(lambda: m4.predict_std(X[16_384:]))
    n_cpu = 1024
    Xt_cpu = m4.X_[:n_cpu]
    _, t_cpu = cpu_timed(lambda: orc.dual_fit(Xt_cpu, y[:n_cpu].astype(np.float64), np.ones(n_cpu), False))
    out["c4"] = {"fit_s": t_fit, "fit_first_call_s": t_fit_first, "gamma_index": int(np.argmin(np.abs(m4.γs_ - m4.γ_))), "predict_std_2k_s": t_std,
                 "eigensolver": "hand-written tridiagonalisation + divide and conquer (csrc/hetrd.cuh, csrc/stedc.cuh)",
                 "cpu_reference_n1024_s": t_cpu,
                 "cpu_sample": "einsum-free oracle port at n = 1024 (the reference needs a 0.27 TB tensor at n = 16,384; its "
                               "cost grows like n^3)"}
    del m4
    torch.cuda.empty_cache()
    # ---- C5-shaped: d = 128, num_features = 4096 (m = 4097) ----
    n5, d5, D5 = 24_000, 128, 4096
    X, y = fast_regression_rows(n5, d5, 64)
    fm = OrthogonalRandomFourierFeatures(num_features=D5).fit(X[:8000], y[:8000], np.ones(8000))
    shift, W = fm.device_weights(d5)
    args5 = (up(X), up(y), up(np.full(n5, 1.0 / n5)), up(shift), up(W))
    _primal.primal_fit(*args5, False, ctx=ctx)
    f5, t_fit = timed(lambda: _primal.primal_fit(*args5, False, ctx=ctx, time_stages=True))
    out["c5_shaped"] = {"rows": n5, "solve_s": t_fit, "stage_ms": f5.stage_ms, "gamma_index": f5.opt,
                        "note": "m = 4097: the eigensolve is the hand-written tridiagonalisation + divide and conquer"}
    # ---- stage 5 at D = 1024: batched predict + predict_std + quantile epilogue over 1M rows ----
    n_s5, d, D = 1_000_000, N_FEATURES, NUM_RFF
    shift, W = bench_map(d, D)
    Xb, yb = fast_regression_rows(N_ROWS, d, N_INFORMATIVE, row_begin=0, row_end=n_s5)
    tr = 65_536
    fit = _primal.primal_fit(up(Xb[:tr]), up(yb[:tr]), up(np.full(tr, 1.0 / tr)), up(shift), up(W), False, ctx=ctx)
    Uinv = ctx.triangular_inverse(fit.U)
    ones = torch.ones(D + 1, dtype=torch.float64, device=dev)
    Xd, shd, Wd = up(Xb), up(shift), up(W)
    q = 3
    cq = [up(np.array([[0.5, 1.0, 1.5], [0.01, 0.0, -0.01], [-1.0, 0.0, 1.0]])), up(np.array([[0.1, 0.2, 0.3], [0.0, 0.0, 0.0], [-0.1, 0.0, 0.1]])),
          up(np.zeros(q)), up(np.zeros(q))]

    def stage5():
        yhat, sigma = ctx.primal_predict(Xd, shd, Wd, beta=fit.beta, B=Uinv, w=ones, want_std=True, b_upper=True)
        return ctx.quantile_epilogue(yhat, sigma, cq[0], cq[1], cq[2], cq[3], True)

    stage5()
    ctx.profile(True)
    _, t_s5 = timed(stage5)
    prof = ctx.profile_read()
    ctx.profile(False)
    m = D + 1
    var_flops = 4.0 * m * m * n_s5  # triangular U^-1: half of the 8 m^2 complex contraction
    var_tf = var_flops / (prof["variance"]["ms"] * 1e-3) / 1e12 if prof["variance"]["ms"] > 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    yhat, sigma = ctx.primal_predict(Xd, shd, Wd, beta=fit.beta, B=Uinv, w=ones, want_std=True, b_upper=True)
    e0.record()
    ctx.quantile_epilogue(yhat, sigma, cq[0], cq[1], cq[2], cq[3], True)
    e1.record()
    torch.cuda.synchronize()
    epi_ms = e0.elapsed_time(e1)
    epi_gbs = n_s5 * (16 + 8 * q) / (epi_ms * 1e-3) / 1e9
    rows_cpu = 20_000
    phi = orc.feature_map(Xb[:rows_cpu], shift.reshape(1, -1), np.ones((1, d)), W)
    L = (fit.U.cpu().numpy(), False)
    _, t_cpu = cpu_timed(lambda: (orc.primal_decision(phi, fit.beta.cpu().numpy()), orc.primal_std(phi, L)))
    out["stage5"] = {
        "rows": n_s5, "seconds": t_s5, "rows_per_s": n_s5 / t_s5,
        "workload": f"predict + predict_std + 3-quantile conformal epilogue over {n_s5} rows, d={d}, num_features={D}",
        "roofline_variance": (
            {"bound": "tensor", "kernel": "oz::gemm_kernel_i8<EpiVariance> (4 m^2 flop/row: U^-1 is triangular, so every column "
                                          "tile contracts over its own k range only; 28 INT8 multiply-adds per FP64 one)",
             "achieved": 28.0 * var_tf, "peak": i8_peak, "unit": "TOP/s (INT8)", "frac": 28.0 * var_tf / i8_peak if i8_peak else None,
             "fp64_equivalent_tflops": var_tf, "vs_fp64_dmma_peak": var_tf / peak_tflops if peak_tflops else None}
            if int8 and var_tf else
            {"bound": "tensor", "kernel": "gemm_kernel<MODE_COMPLEX, OpVariance> (4 m^2 flop/row, U^-1 basis)",
             "achieved": var_tf, "peak": peak_tflops, "unit": "TFLOP/s",
             "frac": var_tf / peak_tflops if var_tf and peak_tflops else None}),
        "roofline_quantile_epilogue": {"bound": "hbm", "kernel": "quantile_epilogue_kernel (16 B in + 8 Q B out per row)",
                                       "achieved": epi_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": epi_gbs / hbm_gbs,
                                       "ms": epi_ms},
        "kernel_ms": {k: v["ms"] for k, v in prof.items()},
        "cpu_reference_rows_per_s": rows_cpu / t_cpu,
        "cpu_sample": f"decision_function + predict_std (n-right-hand-side cho_solve) on {rows_cpu} rows, oracle port",
    }
    return out


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from neo_ls_svm_b200 import _lib, _primal
    from neo_ls_svm_b200.datasets import fast_regression_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, D = args.rows, N_FEATURES, NUM_RFF
    r0, r1 = rank * n // world, (rank + 1) * n // world
    shift, W = bench_map(d, D)  # the reference's own fitted map, shared with the reference arm
    X, y = fast_regression_rows(n, d, N_INFORMATIVE, row_begin=r0, row_end=r1)
    s = np.full(r1 - r0, 1.0 / n)  # uniform weights normalised by the global sum (:110)
    # Pinned host copies for the end-to-end measurement.
    Xh = torch.from_numpy(X).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    sh = torch.from_numpy(s).pin_memory()
    shd, Wd = torch.from_numpy(shift).to(dev), torch.from_numpy(W).to(dev)
    ctx = _lib.context(local_rank)
    ctx.set_eigensolver(args.eig)
    # FP64 tensor peak of this device, measured before any sustained load ("burst" figure).
    peak_burst = ctx.dmma_peak_tflops(20000) if rank == 0 else 0.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(Xd, yd, sd):
        return _primal.primal_fit(Xd, yd, sd, shd, Wd, classifier=False, n_global=n, ctx=ctx)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ----------------------------------------------------------------
    Xd, yd, sd = Xh.to(dev), yh.to(dev), sh.to(dev)
    for _ in range(args.warmup):
        fit = solve(Xd, yd, sd)
    peak_sustained = ctx.dmma_peak_tflops(20000) if rank == 0 else 0.0  # same loop right after the warm-up fits
    peak_tflops = max(peak_burst, peak_sustained)
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: solve(Xd, yd, sd), args.steps)  # per-launch event profiling is OFF in the timed region
    clocks = sampler.stop() if rank == 0 else {}
    launches = torch.tensor([ctx.launch_count() - launches0], dtype=torch.float64, device=dev)
    # Per-kernel device times for the roofline: a separate pass over the same K steps with CUDA events around every
    # GEMM launch (they cost ~1% of a step, which is why they are kept out of `value`).
    ctx.profile(True)
    timed(lambda: solve(Xd, yd, sd), args.steps)
    prof = ctx.profile_read()
    ctx.profile(False)
    # INT8 peaks, measured AFTER the timed region (half a second of full-rate tensor load on rank 0 right before it would
    # heat that GPU and skew the max-over-ranks time).  The INT8 stages run inside a long step at the board's power cap, so
    # their denominator is the SUSTAINED rate of the resident-tile loop: ~0.5 s of back-to-back launches, the lowest of the
    # last five readings; the burst reading is the first call.
    i8_peak_burst = ctx.i8_peak_tops(20000, 256) if rank == 0 else 0.0
    i8_peak_n64 = ctx.i8_peak_tops(20000, 64) if rank == 0 else 0.0
    i8_peak_sustained = min([ctx.i8_peak_tops(20000, 256) for _ in range(20)][-5:]) if rank == 0 else 0.0
    i8_peak = i8_peak_sustained
    if world > 1:
        dist.all_reduce(launches)
    value = n * args.steps / (ms_total * 1e-3)

    # ---- end-to-end arm: host buffers in, host results out, every step ---------------------------
    out_host = {}

    rows_host = torch.empty((5, r1 - r0), dtype=torch.float64).pin_memory()

    def e2e_step():
        # Host rows in through the C ABI (nls_primal_gram_h2d uploads them underneath the Gram pass), every per-row
        # result and β̂ back out to host memory.
        f = _primal.primal_fit(None, None, None, shd, Wd, classifier=False, n_global=n, ctx=ctx, host_rows=(Xh, yh, sh))
        rows_host.copy_(f.rows["_stacked"], non_blocking=True)
        out_host["beta"] = f.beta.cpu()
        torch.cuda.synchronize()
        out_host["rows"] = rows_host

    del Xd, yd, sd
    torch.cuda.empty_cache()
    e2e_step()  # warm-up of the copy path
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = n * args.steps / (ms_e2e * 1e-3)
    h2d = (Xh.numel() + yh.numel() + sh.numel()) * 8 * world
    x_shard_mb = Xh.numel() * 8 / 1e6
    d2h = ((D + 1) * 16 + 1024 * 3 * 8 + 5 * (r1 - r0) * 8) * world

    # ---- public-API arm (N = 1 only): NeoLSSVM(...).fit(X, y) on host NumPy arrays, everything included ----
    fit_api = None
    if world == 1 and not args.skip_api:
        from neo_ls_svm_b200 import NeoLSSVM, OrthogonalRandomFourierFeatures

        del Xh, yh, sh

        def api_fit(rows):
            est = NeoLSSVM(primal_feature_map=OrthogonalRandomFourierFeatures(num_features=D), dual=False)
            t0 = time.perf_counter()
            est.fit(X[:rows], y[:rows])
            torch.cuda.synchronize()
            return est, time.perf_counter() - t0

        api_fit(min(n, 50_000))  # warm-up: numba JIT of the host pre-pass
        est, secs_first = api_fit(n)  # first full-size call of the process: also allocates ~40 GB of device scratch
        del est
        est, secs = api_fit(n)        # steady state, like the W warm-up steps of the solve arm
        phases = dict(getattr(est, "fit_phases_", {}))
        fit_api = {"value": n / secs, "unit": "rows/s", "seconds": secs, "first_call_seconds": secs_first,
                   "phases_s": phases, "selected_gamma_index":
                   int(np.argmin(np.abs(est.γs_ - est.γ_))),
                   "includes": "validation, supervised affine pre-pass (host + GPU weighted-median kernels), ORF, "
                               "pageable H2D, stages 1-4c, D2H, conformal split"}
        del est

    # ---- the other configurations and stage 5 (N = 1 only) ----
    configs = None
    if world == 1 and not args.skip_configs:
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                hbm_gbs = float(json.load(fh)["hbm_gbs"])
        except Exception:  # noqa: BLE001
            hbm_gbs = 6545.6  # the pool's measured copy bandwidth (B200_PROFILING.md fallback)
        try:
            configs = other_configs(ctx, peak_tflops, hbm_gbs, i8_peak)
        except Exception as exc:  # noqa: BLE001  (never lose the headline line to a side measurement)
            configs = {"error": repr(exc)}

    # ---- size-independent correctness properties of the full-size result (outside every timed region) ----
    def rmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lev = fit.rows["loo_leverage"]
    checks = {
        # A is Hermitian; β̂ solves (γ*C + A) β̂ = b; per row loo_residual (1 − leverage) = residual (different kernels)
        "gram_hermitian_rel": float((fit.A - fit.A.conj().T).abs().max() / fit.A.abs().max()),
        "normal_equations_rel": float((fit.A @ fit.beta + (fit.gamma / fit.inv_c) * fit.beta - fit.b).abs().max() / fit.b.abs().max()),
        "loo_identity_rel": rmax(((fit.rows["loo_residuals"] * (1.0 - lev) - fit.rows["residuals"]).abs().max()
                                  / fit.rows["residuals"].abs().max()).item()),
        "leverage_max": rmax(lev.max().item()),
        # the end-to-end arm (host rows in, host results out) reproduces the device-resident arm bit for bit
        "e2e_rows_bitwise_equal": rmax(0.0 if torch.equal(out_host["rows"], fit.rows["_stacked"].cpu()) else 1.0) == 0.0,
    }

    if rank == 0:
        m = D + 1
        rows_local = r1 - r0
        fit_tflops = n * flops_per_row(d, D, N_GAMMAS) * args.steps / (ms_total * 1e-3) / 1e12
        kernel_share = {k: v["ms"] for k, v in prof.items()}
        int8 = os.environ.get("NLS_GEMM", "ozaki") != "dmma"
        int8_sweep = os.environ.get("NLS_GEMM", "ozaki") == "ozaki"
        # The three GEMM stages, each against the peak of the tensor path it runs on.  Algorithmic work (SURVEY.md §8d /
        # DESIGN.md §4): Gram 4m², projection 8m², sweep 4mG FP64 flop per row; on the INT8 core one FP64 multiply-add is
        # 28 exact INT8 digit-plane multiply-adds (7 planes per operand, levels p + q <= 6), so its INT8 work is 28 x that.
        # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch (one 32,768-row chunk) from the `ncu --set full`
        # captures summarised in profiles/r2q_ncu_full_int8.md (final INT8 Gram / projection; the projection writes the sweep's
        # operand planes: 550 MB in, 915 MB out), profiles/r2g_ncu_full_int8.md (first INT8 version, DMMA sweep) and
        # profiles/r1_ncu_full_gemm_kernels.md (DMMA kernels).
        stage_defs = {
            "gram": (4.0 * m * m, int8, "oz::gemm_kernel_i8<EpiGram>" if int8 else "gemm_kernel<MODE_COMPLEX, OpGram>",
                     1.8838e9 if int8 else None),
            "project": (8.0 * m * m, int8, "oz::gemm_kernel_i8<EpiProject> (T = φQ)" if int8 else "gemm_kernel<MODE_COMPLEX, OpProject> (T = φQ)",
                        (1.4641e9 if int8_sweep else 1.0617e9) if int8 else 1.0895e9),
            "sweep": (4.0 * m * N_GAMMAS, int8_sweep,
                      "oz::gemm_kernel_i8<EpiSweep> (fused LOO residual / reduction)" if int8_sweep
                      else "gemm_kernel<MODE_DUAL_A, OpSweep> (fused LOO residual / reduction)", 1.2755e9 if int8_sweep else 0.8093e9),
        }
        stages = {}
        for name, (fpr, on_int8, kernel, traffic) in stage_defs.items():
            ms = prof[name]["ms"]
            tf = fpr * rows_local * args.steps / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            if on_int8:
                stages[name] = {"bound": "tensor", "kernel": kernel, "achieved": 28.0 * tf, "peak": i8_peak, "unit": "TOP/s (INT8)",
                                "frac": 28.0 * tf / i8_peak if i8_peak else None, "fp64_equivalent_tflops": tf,
                                "vs_fp64_dmma_peak": tf / peak_tflops if peak_tflops else None,
                                "frac_of_tile_shape_peak": 28.0 * tf / i8_peak_n64 if i8_peak_n64 else None}
            else:
                stages[name] = {"bound": "tensor", "kernel": kernel, "achieved": tf, "peak": peak_tflops, "unit": "TFLOP/s",
                                "frac": tf / peak_tflops if peak_tflops else None}
            stages[name]["ms_per_step"] = ms / args.steps
            stages[name]["traffic"] = traffic if rows_local >= 32768 else None
        dominant = max(stages, key=lambda k: stages[k]["ms_per_step"])
        cpu = None
        if world == 1:
            v, threads, kind, detail = cpu_reference_rows_per_s(args.cpu_rows, 1, 1)
            cpu = {"value": v, "unit": "rows/s", "cores": threads, "kind": kind, "host_cpus": os.cpu_count(),
                   "sample": f"first {args.cpu_rows} rows of the same dataset, feature-map transform + _optimize_β̂_γ, "
                             f"1 timed pass after 1 warm-up; {detail}"}
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"C3 primal fit n={n} d={d} m={D} G={N_GAMMAS}, rows sharded over {world} GPU(s)",
                "rows_per_gpu": rows_local, "chunk_rows": int(os.environ.get("NLS_CHUNK_ROWS", 32768)),
                "cache": "inputs larger than L2 (X shard %.0f MB; every chunk's feature/projection buffers "
                         "stream through HBM)" % x_shard_mb,
                "feature_map": f"OrthogonalRandomFourierFeatures({D}) fitted by the reference on the first {PREPASS_ROWS} rows "
                               "(tests/golden/bench_c3_map.npz; both arms use it)",
                "selected_gamma_index": fit.opt,
                "eigensolver": args.eig, "jacobi_sweeps": ctx.last_eig_sweeps(),
            },
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches.item()),
            "clocks": clocks,
            "roofline": {
                **{k: v for k, v in stages[dominant].items() if k != "ms_per_step"},
                "stage": dominant,
                "peak_source": "FP64: DMMA register-resident loop measured in this run (nls_bench_dmma_peak), the larger of a "
                               "cold-start burst and a post-warm-up reading (148 SMs x 64 FMA/clk x 1.965 GHz = 37.2 TFLOP/s "
                               "nominal). INT8: resident-tile tcgen05.mma kind::i8 loop measured in this run "
                               "(nls_bench_i8_peak, M = 128, N = 256; 148 SMs x 8192 MAC/clk = 4.76 POP/s nominal at 1.965 GHz), the "
                               "sustained reading after 0.5 s of back-to-back launches (the INT8 stages run at the power cap "
                               "inside a long step; the burst reading is reported next to it). "
                               "MEASURED_PEAKS.json has neither figure",
                "peak_fp64_dmma": peak_tflops, "peak_burst": peak_burst, "peak_sustained": peak_sustained,
                "peak_int8": i8_peak, "peak_int8_burst": i8_peak_burst, "peak_int8_sustained": i8_peak_sustained,
                "peak_int8_at_tile_shape_n64": i8_peak_n64,
                "stages": stages,
                "fit_tflops": fit_tflops, "fit_frac": fit_tflops / (peak_tflops * world) if peak_tflops else None,
                "fit_frac_note": "whole-fit algorithmic FP64 flop rate over the FP64 DMMA peak; above 1 means the INT8 "
                                 "(Ozaki) stages beat what FP64 tensor cores could do at 100 %",
                "kernel_ms": kernel_share,
            },
            "cpu_baseline": cpu,
            "fit_api": fit_api,
            "configs": configs,
            "checks": checks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rows", type=int, default=N_ROWS, help="total training rows (default: config C3)")
    ap.add_argument("--cpu-rows", type=int, default=CPU_SAMPLE_ROWS,
                    help="rows per step of the CPU reference measurement (BASELINE.md §3: 100,000)")
    ap.add_argument("--skip-configs", action="store_true", help="skip the C1/C2/C4/C5-shaped/stage-5 timings (N=1 arm)")
    ap.add_argument("--skip-api", action="store_true", help="skip the public-API NeoLSSVM.fit timing (N=1 arm)")
    ap.add_argument("--eig", choices=["auto", "dc", "jacobi", "cusolver"], default="auto",
                    help="stage-3 eigensolver: auto = dc = hand-written tridiagonalisation + divide and conquer; jacobi = hand-written block "
                         "Jacobi; cusolver = library comparator")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
