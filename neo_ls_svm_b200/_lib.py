"""ctypes binding of libnls_b200.so (the C ABI declared in include/nls_b200.h).

PyTorch supplies device memory, the stream and (for multi-GPU) `torch.distributed`; every numerical
stage of the hot path runs in the CUDA library.  There is no CPU fallback: if the library is missing
or no sm_100 GPU is present, the first call raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnls_b200.so")

PROF_KINDS = ("feature_map", "gram", "project", "sweep", "variance", "other", "slice")

_c_double_p = C.c_void_p  # device pointers are passed as integers
_lib = None
_lib_lock = threading.Lock()


class NlsError(RuntimeError):
    """Raised when a libnls_b200 entry point returns a non-zero status."""


def _signatures(lib: C.CDLL) -> None:
    p, i, i64, d = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "nls_version": ([], i),
        "nls_last_error": ([], C.c_char_p),
        "nls_ctx_create": ([i, p, C.POINTER(p)], i),
        "nls_ctx_destroy": ([p], i),
        "nls_ctx_set_stream": ([p, p], i),
        "nls_ctx_set_chunk_rows": ([p, i64], i),
        "nls_ctx_launch_count": ([p], i64),
        "nls_ctx_profile": ([p, i], i),
        "nls_ctx_profile_read": ([p, C.POINTER(d), C.POINTER(i64)], i),
        "nls_feature_map": ([p, p, i64, i, p, p, i, p], i),
        "nls_affine_map": ([p, p, i64, i, p, p, i, p], i),
        "nls_primal_gram": ([p, p, p, p, i64, i, p, p, i, p, p], i),
        "nls_primal_gram_h2d": ([p, p, p, p, i64, i, p, p, p, p, p, i, p, p], i),
        "nls_heev": ([p, p, i, d, p, p], i),
        "nls_stedc": ([p, i, p, p, p, p], i),
        "nls_ctx_last_tridiagonal": ([p, i, p, p], i),
        "nls_ctx_set_eigensolver": ([p, i], i),
        "nls_ctx_set_gemm_core": ([p, i], i),
        "nls_ctx_last_eig_sweeps": ([p], i),
        "nls_primal_coeffs": ([p, p, p, p, i, d, d, p, p], i),
        "nls_cholesky_solve": ([p, p, i, d, p, p, p], i),
        "nls_triangular_inverse": ([p, p, i, i, p], i),
        "nls_primal_loo_sweep": ([p, p, p, p, i64, i, p, p, i, p, p, p, d, p, i, i, p, p], i),
        "nls_primal_finalize": ([p, p, p, p, i64, i, p, p, i, p, p, d, d, p, p, i, p, p, p, p, p, p], i),
        "nls_primal_predict": ([p, p, i64, i, p, p, i, p, p, p, i, p, p], i),
        "nls_quantile_epilogue": ([p, p, p, i64, p, p, p, p, i, i, p, p, i, p], i),
        "nls_dual_sweep": ([p, p, i, i, p, p, p, p, i, i, p, p, p], i),
        "nls_dual_finalize": ([p, i, p, p, d, p, p, p, p, p, p, p], i),
        "nls_dual_predict": ([p, p, i64, p, i, i, p, d, p, p, p, p], i),
        "nls_bin_median_stats": ([p, p, i64, i, p, p, p, i, p, i, p, p, p], i),
        "nls_bin_mad": ([p, p, i64, i, p, p, p, i, p, i, p, p], i),
        "nls_bench_dmma_peak": ([p, i, C.POINTER(d)], i),
        "nls_bench_i8_peak": ([p, i, i, C.POINTER(d)], i),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype


EXPORTED = None  # filled by load(): names bound above


def load() -> C.CDLL:
    """Load libnls_b200.so (building it first if nvcc is available and it is missing)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import _build

            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                raise NlsError(
                    f"{LIB_PATH} is missing and could not be built ({exc}). The Neo LS-SVM hot path has "
                    "no CPU fallback: build it with `python -m neo_ls_svm_b200._build`."
                ) from exc
        lib = C.CDLL(LIB_PATH)
        _signatures(lib)
        _lib = lib
        return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().nls_last_error()
        raise NlsError(f"libnls_b200 error {status}: {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a contiguous float64/complex128 CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    import torch

    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    assert t.dtype in (torch.float64, torch.complex128, torch.int64, torch.int32), f"unexpected dtype {t.dtype}"
    return t.data_ptr()


class Context:
    """One library context per (process, device), bound to torch's current stream at creation."""

    def __init__(self, device: int | None = None):
        import torch

        if not torch.cuda.is_available():
            raise NlsError("No CUDA device: the Neo LS-SVM B200 hot path has no CPU fallback.")
        self.lib = load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream().cuda_stream
        handle = C.c_void_p()
        check(self.lib.nls_ctx_create(self.device, C.c_void_p(self.stream), C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.nls_ctx_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass

    # -- bookkeeping ---------------------------------------------------------------------------
    def bind_current_stream(self) -> None:
        """Enqueue on torch's CURRENT stream of this device from now on (the tensors handed to the library are
        produced and consumed on that stream).  Called by `context()` on every lookup."""
        import torch

        stream = torch.cuda.current_stream(self.device).cuda_stream
        if stream != self.stream:
            check(self.lib.nls_ctx_set_stream(self.handle, C.c_void_p(stream)))
            self.stream = stream

    def set_chunk_rows(self, rows: int) -> None:
        check(self.lib.nls_ctx_set_chunk_rows(self.handle, rows))

    def set_eigensolver(self, kind: str) -> None:
        """'dc' (hand-written tridiagonalisation + divide and conquer; what 'auto', the default, runs), 'jacobi'
        (hand-written block Jacobi kernels) or 'cusolver' (library comparator)."""
        check(self.lib.nls_ctx_set_eigensolver(self.handle, {"jacobi": 0, "cusolver": 1, "auto": 2, "dc": 3}[kind]))

    def set_gemm_core(self, kind: str) -> None:
        """'ozaki' (default): Gram, projection, γ sweep and predict_std run on the INT8 tensor cores (tcgen05, Ozaki scheme,
        FP64-accurate); 'ozaki-dmma-sweep': the sweep stays on FP64 DMMA; 'dmma': FP64 DMMA everywhere."""
        check(self.lib.nls_ctx_set_gemm_core(self.handle, {"dmma": 0, "ozaki-dmma-sweep": 1, "ozaki": 2}[kind]))

    def last_eig_sweeps(self) -> int:
        return int(self.lib.nls_ctx_last_eig_sweeps(self.handle))

    def launch_count(self) -> int:
        return int(self.lib.nls_ctx_launch_count(self.handle))

    def profile(self, enable: bool) -> None:
        check(self.lib.nls_ctx_profile(self.handle, int(enable)))

    def profile_read(self) -> dict:
        n = len(PROF_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        check(self.lib.nls_ctx_profile_read(self.handle, ms, cnt))
        return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(PROF_KINDS)}

    def dmma_peak_tflops(self, iters: int = 20000) -> float:
        out = C.c_double()
        check(self.lib.nls_bench_dmma_peak(self.handle, iters, C.byref(out)))
        return out.value

    def i8_peak_tops(self, iters: int = 20000, n_cols: int = 256) -> float:
        """INT8 tensor-core rate (TOP/s) of a resident-tile tcgen05.mma kind::i8 loop: n_cols = 256 is the hardware peak,
        64 the shape the Ozaki tile uses."""
        out = C.c_double()
        check(self.lib.nls_bench_i8_peak(self.handle, iters, n_cols, C.byref(out)))
        return out.value

    # -- stages --------------------------------------------------------------------------------
    def feature_map(self, X, shift, W):
        import torch

        n, d = X.shape
        D = W.shape[1]
        phi = torch.empty((n, D + 1), dtype=torch.complex128, device=X.device)
        check(self.lib.nls_feature_map(self.handle, ptr(X), n, d, ptr(shift), ptr(W), D, ptr(phi)))
        return phi

    def affine_map(self, X, shift, W):
        import torch

        n, d = X.shape
        D = W.shape[1]
        Z = torch.empty((n, D), dtype=torch.float64, device=X.device)
        check(self.lib.nls_affine_map(self.handle, ptr(X), n, d, ptr(shift), ptr(W), D, ptr(Z)))
        return Z

    def primal_gram(self, X, y, s, shift, W):
        import torch

        n, d = X.shape
        D = W.shape[1]
        m = D + 1
        A = torch.empty((m, m), dtype=torch.complex128, device=X.device)
        b = torch.empty((m,), dtype=torch.complex128, device=X.device)
        check(self.lib.nls_primal_gram(self.handle, ptr(X), ptr(y), ptr(s), n, d, ptr(shift), ptr(W), D, ptr(A), ptr(b)))
        return A, b

    def primal_gram_h2d(self, Xh, yh, sh, shift, W):
        """Stage 1+2 from HOST tensors (contiguous float64 CPU tensors, ideally pinned): the rows are uploaded in
        groups on a copy stream underneath the Gram pass.  Returns (A, b, X, y, s) with the device copies."""
        import torch

        n, d = Xh.shape
        D = W.shape[1]
        m = D + 1
        dev = W.device
        for t in (Xh, yh, sh):
            assert t.device.type == "cpu" and t.dtype == torch.float64 and t.is_contiguous()
        X = torch.empty((n, d), dtype=torch.float64, device=dev)
        y = torch.empty((n,), dtype=torch.float64, device=dev)
        s = torch.empty((n,), dtype=torch.float64, device=dev)
        A = torch.empty((m, m), dtype=torch.complex128, device=dev)
        b = torch.empty((m,), dtype=torch.complex128, device=dev)
        check(self.lib.nls_primal_gram_h2d(self.handle, Xh.data_ptr(), yh.data_ptr(), sh.data_ptr(), n, d, ptr(X), ptr(y), ptr(s),
                                           ptr(shift), ptr(W), D, ptr(A), ptr(b)))
        return A, b, X, y, s

    def heev(self, A, scale: float):
        import torch

        m = A.shape[0]
        lam = torch.empty((m,), dtype=torch.float64, device=A.device)
        Q = torch.empty((m, m), dtype=torch.complex128, device=A.device)
        check(self.lib.nls_heev(self.handle, ptr(A), m, float(scale), ptr(lam), ptr(Q)))
        return lam, Q

    def stedc(self, d, e):
        """Eigenpairs of the real symmetric tridiagonal matrix (d, e: host NumPy arrays): (lam, Zt) on the device,
        row k of Zt the eigenvector of lam[k]."""
        import numpy as np
        import torch

        d = np.ascontiguousarray(d, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        n = len(d)
        dev = torch.device("cuda", self.device)
        lam = torch.empty((n,), dtype=torch.float64, device=dev)
        Zt = torch.empty((n, n), dtype=torch.float64, device=dev)
        check(self.lib.nls_stedc(self.handle, n, d.ctypes.data, e.ctypes.data if n > 1 else None, ptr(lam), ptr(Zt)))
        return lam, Zt

    def last_tridiagonal(self, n: int):
        import numpy as np

        d, e = np.empty(n), np.empty(max(n - 1, 0))
        check(self.lib.nls_ctx_last_tridiagonal(self.handle, n, d.ctypes.data, e.ctypes.data if n > 1 else None))
        return d, e

    def primal_coeffs(self, Q, lam, b, inv_c: float, gamma: float | None = None, v=None):
        """v = Q^H b inv_c (if b is given) and beta_eig = Q (v/(lam+gamma)) (if gamma is given)."""
        import torch

        m = Q.shape[0]
        if v is None:
            v = torch.empty((m,), dtype=torch.complex128, device=Q.device)
        beta = torch.empty((m,), dtype=torch.complex128, device=Q.device) if gamma is not None else None
        check(self.lib.nls_primal_coeffs(
            self.handle, ptr(Q), ptr(lam), ptr(b), m, float(inv_c), float(gamma if gamma is not None else 0.0),
            ptr(v), ptr(beta)))
        return v, beta

    def cholesky_solve(self, A, diag_shift: float, b=None):
        import torch

        m = A.shape[0]
        U = torch.empty((m, m), dtype=torch.complex128, device=A.device)
        beta = torch.empty((m,), dtype=torch.complex128, device=A.device) if b is not None else None
        check(self.lib.nls_cholesky_solve(self.handle, ptr(A), m, float(diag_shift), ptr(b), ptr(U), ptr(beta)))
        return U, beta

    def triangular_inverse(self, U):
        """U⁻¹ of the upper-triangular factor U (anything below the diagonal of U is ignored): the basis B of
        predict_std, (γC + A)⁻¹ = U⁻¹ U⁻ᴴ (_neo_ls_svm.py:467-469, :473-475)."""
        import torch

        m = U.shape[0]
        U = U.contiguous()
        out = torch.empty((m, m), dtype=U.dtype, device=U.device)
        check(self.lib.nls_triangular_inverse(self.handle, ptr(U), m, int(U.is_complex()), ptr(out)))
        return out

    def primal_loo_sweep(self, X, y, s, shift, W, Q, lam, v, inv_c: float, gammas, classifier: bool, stash=None):
        """Per-γ error sums (3×G); `stash` (n×G float64, optional) receives σ²ᵢ(γ_g) for every row."""
        import torch

        n, d = X.shape
        D = W.shape[1]
        G = gammas.shape[0]
        sums = torch.empty((3, G), dtype=torch.float64, device=X.device)
        if stash is not None:
            assert stash.shape == (n, G)
        check(self.lib.nls_primal_loo_sweep(
            self.handle, ptr(X), ptr(y), ptr(s), n, d, ptr(shift), ptr(W), D, ptr(Q), ptr(lam), ptr(v),
            float(inv_c), ptr(gammas), G, int(classifier), ptr(sums), ptr(stash)))
        return sums

    def primal_finalize(self, X, y, s, shift, W, Q, lam, inv_c: float, gamma: float, beta_eig, beta, classifier: bool,
                        sigma2=None):
        import torch

        n, d = X.shape
        D = W.shape[1]
        out = torch.empty((5, n), dtype=torch.float64, device=X.device)
        check(self.lib.nls_primal_finalize(
            self.handle, ptr(X), ptr(y), ptr(s), n, d, ptr(shift), ptr(W), D, ptr(Q), ptr(lam), float(inv_c),
            float(gamma), ptr(beta_eig), ptr(beta), int(classifier), ptr(sigma2),
            ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]), ptr(out[4])))
        return {"loo_residuals": out[0], "yhat_loo": out[1], "loo_leverage": out[2], "residuals": out[3],
                "loo_std": out[4], "_stacked": out}

    def primal_predict(self, X, shift, W, beta=None, B=None, w=None, want_std: bool = False, b_upper: bool = False):
        import torch

        n, d = X.shape
        D = W.shape[1]
        yhat = torch.empty((n,), dtype=torch.float64, device=X.device) if beta is not None else None
        sigma = torch.empty((n,), dtype=torch.float64, device=X.device) if want_std else None
        check(self.lib.nls_primal_predict(
            self.handle, ptr(X), n, d, ptr(shift), ptr(W), D, ptr(beta), ptr(B), ptr(w), int(b_upper), ptr(yhat), ptr(sigma)))
        return yhat, sigma

    def quantile_epilogue(self, yhat, sigma, beta_abs, beta_rel, bias_abs, bias_rel, regressor: bool,
                          iso_x=None, iso_y=None):
        import torch

        n = yhat.shape[0]
        Q = bias_abs.shape[0]
        shape = (n, Q) if regressor else (n, Q, 2)
        out = torch.empty(shape, dtype=torch.float64, device=yhat.device)
        n_iso = 0 if iso_x is None else iso_x.shape[0]
        check(self.lib.nls_quantile_epilogue(
            self.handle, ptr(yhat), ptr(sigma), n, ptr(beta_abs), ptr(beta_rel), ptr(bias_abs), ptr(bias_rel),
            Q, int(regressor), ptr(iso_x), ptr(iso_y), n_iso, ptr(out)))
        return out


    # -- dual path -----------------------------------------------------------------------------
    def dual_sweep(self, Xt, y, s, sn, gammas, classifier: bool):
        import torch

        n, p_ = Xt.shape
        G = gammas.shape[0]
        sums = torch.empty((3, G), dtype=torch.float64, device=Xt.device)
        yhat_loo = torch.empty((n, G), dtype=torch.float64, device=Xt.device)
        lam = torch.empty((n,), dtype=torch.float64, device=Xt.device)
        check(self.lib.nls_dual_sweep(
            self.handle, ptr(Xt), n, p_, ptr(y), ptr(s), ptr(sn), ptr(gammas), G, int(classifier),
            ptr(sums), ptr(yhat_loo), ptr(lam)))
        return sums, yhat_loo, lam

    def dual_finalize(self, n: int, y, sn, gamma: float, cholesky: bool = True):
        import torch

        dev = y.device
        f64 = torch.float64
        out = {
            "alpha": torch.empty((n,), dtype=f64, device=dev),
            "alpha_eig": torch.empty((n,), dtype=f64, device=dev),
            "U": torch.empty((n, n), dtype=f64, device=dev) if cholesky else None,
            "Falpha": torch.empty((n,), dtype=f64, device=dev),
            "sigma2": torch.empty((n,), dtype=f64, device=dev),
            "Bt": torch.empty((n, n), dtype=f64, device=dev),
            "w": torch.empty((n,), dtype=f64, device=dev),
        }
        check(self.lib.nls_dual_finalize(
            self.handle, n, ptr(y), ptr(sn), float(gamma), ptr(out["alpha"]), ptr(out["alpha_eig"]), ptr(out["U"]),
            ptr(out["Falpha"]), ptr(out["sigma2"]), ptr(out["Bt"]), ptr(out["w"])))
        return out

    def dual_predict(self, Xq, Xt, alpha=None, alpha_sum: float = 0.0, Bt=None, w=None, want_std: bool = False):
        import torch

        nq, p_ = Xq.shape
        n = Xt.shape[0]
        if Xt.shape[1] != p_:
            raise ValueError(f"query rows have {p_} transformed features, the training rows {Xt.shape[1]}")
        yhat = torch.empty((nq,), dtype=torch.float64, device=Xq.device) if alpha is not None else None
        sigma = torch.empty((nq,), dtype=torch.float64, device=Xq.device) if want_std else None
        check(self.lib.nls_dual_predict(
            self.handle, ptr(Xq), nq, ptr(Xt), n, p_, ptr(alpha), float(alpha_sum), ptr(Bt), ptr(w), ptr(yhat), ptr(sigma)))
        return yhat, sigma


    # -- supervised affine pre-pass ------------------------------------------------------------
    def bin_median_stats(self, X, perm, w, tiles, bin_tiles, thresh=None):
        import torch

        n, d = X.shape
        nbins, ntiles = bin_tiles.shape[0], tiles.shape[0]
        stats = torch.empty((7, nbins, d), dtype=torch.float64, device=X.device)
        wtot = torch.empty((nbins, d), dtype=torch.float64, device=X.device)
        check(self.lib.nls_bin_median_stats(
            self.handle, ptr(X), n, d, ptr(perm), ptr(w), ptr(tiles), ntiles, ptr(bin_tiles), nbins, ptr(thresh), ptr(stats), ptr(wtot)))
        return stats, wtot

    def bin_mad(self, X, perm, w, tiles, bin_tiles, centre):
        import torch

        n, d = X.shape
        nbins, ntiles = bin_tiles.shape[0], tiles.shape[0]
        spread = torch.empty((nbins, d), dtype=torch.float64, device=X.device)
        check(self.lib.nls_bin_mad(
            self.handle, ptr(X), n, d, ptr(perm), ptr(w), ptr(tiles), ntiles, ptr(bin_tiles), nbins, ptr(centre), ptr(spread)))
        return spread


_contexts: dict = {}


def context(device: int | None = None) -> Context:
    """Process-wide context for `device` (default: torch's current device)."""
    import torch

    dev = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else device
    if dev not in _contexts:
        _contexts[dev] = Context(dev)
    _contexts[dev].bind_current_stream()
    return _contexts[dev]
