/* nls_b200.h — C ABI of the B200-native Neo LS-SVM hot path (libnls_b200.so).
 *
 * The reference (lsorber/neo-ls-svm v0.3.0) is pure Python and has no FFI layer; the boundary it
 * exposes for this path is the sklearn estimator `NeoLSSVM` (src/neo_ls_svm/_neo_ls_svm.py:43).
 * These entry points are what a maintainer of the reference would bind (ctypes, see
 * INTEGRATION.md) to replace the NumPy/SciPy bodies cited on each function.
 *
 * Conventions
 *   - extern "C"; every function returns 0 on success and a negative nls_status on failure;
 *     nls_last_error() returns a thread-local, human-readable message.  No C++ exceptions cross.
 *   - All array arguments are caller-owned DEVICE pointers on the context's device unless the
 *     function name ends in `_host`.  Matrices are row-major FP64; complex values are interleaved
 *     (re, im) doubles, i.e. NumPy complex128 layout.
 *   - Work is enqueued on the stream given at context creation (torch's current stream in the
 *     Python host layer).  Functions do not synchronise unless they return host scalars.
 *   - One context per process per GPU; multi-GPU runs are one process per GPU and the caller sums
 *     the documented partial results with an NCCL all-reduce between calls.
 *   - There is no CPU fallback: every entry point needs an sm_100a device.
 */
#ifndef NLS_B200_H
#define NLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nls_ctx nls_ctx;

typedef enum {
  NLS_OK = 0,
  NLS_ERR_INVALID = -1, /* bad argument */
  NLS_ERR_CUDA = -2,    /* CUDA runtime / driver error */
  NLS_ERR_ALLOC = -3,   /* device allocation failed */
  NLS_ERR_SOLVER = -4,  /* eigensolver did not converge */
  NLS_ERR_ARCH = -5     /* device is not sm_100 */
} nls_status;

/* Library version (major*10000 + minor*100 + patch). */
int nls_version(void);
/* Thread-local message describing the last failure on this thread. */
const char* nls_last_error(void);

/* Create / destroy a context bound to `device`, enqueueing on `stream` (a cudaStream_t, may be 0). */
int nls_ctx_create(int device, void* stream, nls_ctx** out);
int nls_ctx_destroy(nls_ctx* ctx);
/* Enqueue on another stream from now on (the old stream is drained first: the scratch buffers are shared). */
int nls_ctx_set_stream(nls_ctx* ctx, void* stream);
/* Rows per internal chunk (default 32768); affects scratch size only, never results' meaning. */
int nls_ctx_set_chunk_rows(nls_ctx* ctx, int64_t rows);
/* Kernels of this library launched through `ctx` so far (for bench.py's gpu_launches). */
int64_t nls_ctx_launch_count(const nls_ctx* ctx);
/* Device milliseconds spent in the dominant GEMM kernels since the last reset (CUDA events on the
 * context's stream; only collected when enabled because it adds event records). */
int nls_ctx_profile(nls_ctx* ctx, int enable);
int nls_ctx_profile_read(nls_ctx* ctx, double* ms_out /* [NLS_PROF_N] */, int64_t* launches_out /* [NLS_PROF_N] */);
enum { NLS_PROF_FEATURE_MAP = 0, NLS_PROF_GRAM = 1, NLS_PROF_PROJECT = 2, NLS_PROF_SWEEP = 3,
       NLS_PROF_VARIANCE = 4, NLS_PROF_OTHER = 5, NLS_PROF_SLICE = 6 /* digit-plane slicing of the INT8 core */,
       NLS_PROF_N = 7 };

/* ---------------------------------------------------------------------------------------------
 * Stage 1 — feature map.  Replaces AffineFeatureMap.transform (_affine_feature_map.py:72-92) +
 * RandomFourierFeatures.transform (_feature_maps.py:153-203):
 *     z = (x - shift) W,   phi = [exp(-1j z)/sqrt(D) | 1]
 * X: n x d, shift: d, W: d x D (= A_/scale^T), phi_out: n x (D+1) complex128.
 * ------------------------------------------------------------------------------------------- */
int nls_feature_map(nls_ctx* ctx, const double* X, int64_t n, int d, const double* shift,
                    const double* W, int D, double* phi_out);
/* The affine part alone, z = (x - shift) W as a real n x D matrix: AffineFeatureMap.transform
 * (_affine_feature_map.py:72-92), used by the dual path (_neo_ls_svm.py:394, :473, :668). */
int nls_affine_map(nls_ctx* ctx, const double* X, int64_t n, int d, const double* shift,
                   const double* W, int D, double* Z_out);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 — primal Gram.  Replaces _neo_ls_svm.py:112-114 and :127:
 *     A = (S phi)^H (S phi) (Hermitian, exactly symmetrised),  b = (S phi)^H (s*y)
 * for the local rows.  s must already be normalised by the GLOBAL weight sum (:110).
 * A_out: m x m complex128 (m = D+1), b_out: m complex128.  Partial over rows: all-reduce(sum).
 * ------------------------------------------------------------------------------------------- */
int nls_primal_gram(nls_ctx* ctx, const double* X, const double* y, const double* s, int64_t n,
                    int d, const double* shift, const double* W, int D, double* A_out,
                    double* b_out);
/* The same with the training rows still on the HOST (the reference's fit receives host ndarrays,
 * _neo_ls_svm.py:327-335): X_host (n x d), y_host, s_host are copied into the caller's device buffers
 * X_dev / y_dev / s_dev in row groups on a copy stream while the Gram of the groups that have already
 * arrived is computed, so the upload costs no time; the device copies stay valid for stages 4a-4c.
 * Page-locked host buffers give a truly asynchronous copy; pageable ones are staged by the driver. */
int nls_primal_gram_h2d(nls_ctx* ctx, const double* X_host, const double* y_host,
                        const double* s_host, int64_t n, int d, double* X_dev, double* y_dev,
                        double* s_dev, const double* shift, const double* W, int D, double* A_out,
                        double* b_out);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 — Hermitian eigendecomposition.  Replaces scipy.linalg.eigh at _neo_ls_svm.py:120:
 *     lam, Q = eigh(scale * A)   (ascending lam, Q[:, k] the k-th eigenvector)
 * A: m x m complex128 (only needs to be Hermitian), lam_out: m, Q_out: m x m complex128.
 * ------------------------------------------------------------------------------------------- */
int nls_heev(nls_ctx* ctx, const double* A, int m, double scale, double* lam_out, double* Q_out);
/* The middle stage of nls_heev on its own: eigen-decomposition of the real symmetric tridiagonal matrix with diagonal
 * d_host (n) and off-diagonal e_host (n - 1), both HOST arrays, by divide and conquer (csrc/stedc.cuh).  lam_out
 * (device, n) ascending; Zt_out (device, n x n row-major): row k is the eigenvector of lam_out[k]. */
int nls_stedc(nls_ctx* ctx, int n, const double* d_host, const double* e_host, double* lam_out, double* Zt_out);
/* The tridiagonal form (host arrays d: n, e: n - 1) the last eigensolve of this context reduced its matrix to. */
int nls_ctx_last_tridiagonal(nls_ctx* ctx, int n, double* d_host, double* e_host);
/* Which solver nls_heev (and the dual path's real symmetric solve) runs: 0 = the hand-written parallel two-sided
 * block Jacobi kernels (csrc/jacobi_wide.cuh), 1 = cuSOLVER Zheevd / Dsyevd (library comparator, never the default),
 * 2 = auto (default) = 3 = the hand-written tridiagonalisation + divide and conquer + back-transformation
 * (csrc/hetrd.cuh, csrc/stedc.cuh).  Also selectable with NLS_EIG=jacobi|cusolver|dc. */
int nls_ctx_set_eigensolver(nls_ctx* ctx, int kind);
/* Which tensor-core path the GEMM stages run on.  2 (default) = the Gram (_neo_ls_svm.py:112-114), the eigenbasis
 * projection T = phi Q (:134, :137), the gamma sweep (:147-161) and predict_std (:467-469) on tcgen05.mma kind::i8 with TMEM
 * accumulators, FP64-accurate through the Ozaki scheme (7 balanced base-256 digit planes per operand, 28 exact INT8 plane
 * products, FP64 recombination; csrc/ozaki.cuh); 1 = the same with the sweep on FP64 DMMA; 0 = FP64 DMMA everywhere.
 * Also selectable with NLS_GEMM=ozaki|ozaki-dmma-sweep|dmma. */
int nls_ctx_set_gemm_core(nls_ctx* ctx, int kind);
/* Number of Jacobi sweeps the last nls_heev call needed (0 for cuSOLVER). */
int nls_ctx_last_eig_sweeps(const nls_ctx* ctx);

/* ---------------------------------------------------------------------------------------------
 * Small replicated solves between the stages (O(m^2) / O(m^3), every rank computes the same).
 * ------------------------------------------------------------------------------------------- */
/* v = Q^H b inv_c (_neo_ls_svm.py:121, :129) when b != NULL (else v_out is read as input), and
 * beta_eig = Q (v / (lam + gamma)) (:175) when beta_eig_out != NULL.  All complex128, length m. */
int nls_primal_coeffs(nls_ctx* ctx, const double* Q, const double* lam, const double* b, int m,
                      double inv_c, double gamma, double* v_out, double* beta_eig_out);
/* Replaces cho_factor / cho_solve at _neo_ls_svm.py:177-178: U_out (m x m complex128, upper
 * triangle valid, M = U^H U as scipy.linalg.cho_factor returns it) for M = A + diag_shift * I, and
 * beta_out = M^-1 b when b != NULL. */
int nls_cholesky_solve(nls_ctx* ctx, const double* A, int m, double diag_shift, const double* b,
                       double* U_out, double* beta_out);

/* U^-1 (n x n row-major) of an upper-triangular factor U (n x n row-major; whatever lies below its diagonal is
 * ignored, as with scipy's cho_factor output): the basis predict_std contracts with, (gamma C + A)^-1 = U^-1 U^-H
 * (_neo_ls_svm.py:467-469 primal, :473-475 dual).  is_complex != 0: complex128 (interleaved), else float64. */
int nls_triangular_inverse(nls_ctx* ctx, const double* U, int n, int is_complex, double* B_out);

/* ---------------------------------------------------------------------------------------------
 * Stage 4 — leave-one-out sweep.  Replaces _neo_ls_svm.py:128-165 for the local rows:
 *     T = phi Q;  P = Re(T * v);  H = s^2 |T|^2 inv_c;  loo = (P r - y) / (1 - H r),
 *     r[k, g] = 1 / (gammas[g] + lam[k]);  classifier clip (:153-155);
 *     sums_out[0, g] = sum_i s_i |loo_ig|                        (loo_errors_gammas_, :158)
 *     sums_out[1, g] = sum_i s_i (|loo_ig| >= 1)                  (classifier only, :160)
 *     sums_out[2, g] = sum_i s_i max(0, |loo_ig| - 1)             (classifier only, :161)
 * v = Q^H b inv_c (:121, :129), inv_c = n_global * m (:117-118).  sums_out: 3 x G, partial over
 * rows: all-reduce(sum); the caller takes the argmin (:159-165).
 * sigma2_stash (optional, n x G, may be NULL): receives sigma2_i(gamma_g) = sum_k |T_ik|^2 inv_c r[k,g]
 * for every row and gamma, so that nls_primal_finalize needs no second projection pass (8 KB per
 * row of HBM instead of 8 m^2 flops per row).
 * ------------------------------------------------------------------------------------------- */
int nls_primal_loo_sweep(nls_ctx* ctx, const double* X, const double* y, const double* s,
                         int64_t n, int d, const double* shift, const double* W, int D,
                         const double* Q, const double* lam, const double* v, double inv_c,
                         const double* gammas, int G, int is_classifier, double* sums_out,
                         double* sigma2_stash);

/* ---------------------------------------------------------------------------------------------
 * Stage 4c — per-row outputs at the selected gamma.  Replaces _neo_ls_svm.py:167-187:
 *     sigma2_i   = sum_k |T_ik|^2 inv_c / (lam_k + gamma)      (= phi_i (gamma C + A)^-1 phi_i^H, :184)
 *     leverage_i = s_i^2 sigma2_i                              (:169)
 *     loo_i      = (Re(phi_i beta_eig) - y_i) / (1 - leverage_i), yhat_loo_i = y_i + loo_i (:149-150)
 *     loo_res_i  = classifier-clipped loo_i                    (:153-155, :167)
 *     resid_i    = Re(phi_i beta) - y_i, classifier-clipped    (:179-182)
 *     loo_std_i  = sqrt(sigma2_i + (s_i sigma2_i)^2 / (1 - leverage_i))   (:186-187)
 * beta_eig = Q (v / (lam + gamma)) (:175); beta = Cholesky re-solve (:177-178).  All outputs: n.
 * sigma2_in (optional, n, may be NULL): column `opt` of the stash written by nls_primal_loo_sweep;
 * when given, Q and lam are not read and the projection GEMM is skipped.
 * ------------------------------------------------------------------------------------------- */
int nls_primal_finalize(nls_ctx* ctx, const double* X, const double* y, const double* s, int64_t n,
                        int d, const double* shift, const double* W, int D, const double* Q,
                        const double* lam, double inv_c, double gamma, const double* beta_eig,
                        const double* beta, int is_classifier, const double* sigma2_in,
                        double* loo_res_out,
                        double* yhat_loo_out, double* leverage_out, double* resid_out,
                        double* loo_std_out);

/* ---------------------------------------------------------------------------------------------
 * Stage 5 — batched predict / predict_std.  Replaces decision_function (_neo_ls_svm.py:663-665)
 * and predict_std (:467-469, :477) with one pass over phi(x):
 *     yhat_i = Re(phi_i beta);   sigma_i = sqrt( sum_k |(phi_i B)_k|^2 w_k )
 * B (m x m complex128) and w (m) describe (gamma C + A)^-1 = B diag(w) B^H: either the
 * eigenbasis (B = Q, w = inv_c/(lam+gamma)) or the inverse Cholesky factor (B = U^-1, w = 1; pass
 * b_upper = 1 so that only the non-zero half of the triangular contraction is computed).
 * yhat_out / sigma_out may be NULL to skip that output.
 * ------------------------------------------------------------------------------------------- */
int nls_primal_predict(nls_ctx* ctx, const double* X, int64_t n, int d, const double* shift,
                       const double* W, int D, const double* beta, const double* B,
                       const double* w, int b_upper, double* yhat_out, double* sigma_out);

/* ---------------------------------------------------------------------------------------------
 * Stage 5c — conformal quantile epilogue.  Replaces the per-row part of predict_quantiles
 * (_neo_ls_svm.py:566-600).  beta_abs / beta_rel: F x Q coefficient matrices of the two coherent
 * quantile regressors (F = 3 for a regressor: [sigma, |yhat|, 1]; F = 2 for a classifier:
 * [sigma, 1]); bias_*: Q.  Regressor: out is n x Q.  Classifier: iso_x/iso_y (n_iso isotonic
 * thresholds) are applied per quantile and out is n x Q x 2 ([1 - p reversed | p], :600).
 * ------------------------------------------------------------------------------------------- */
int nls_quantile_epilogue(nls_ctx* ctx, const double* yhat, const double* sigma, int64_t n,
                          const double* beta_abs, const double* beta_rel, const double* bias_abs,
                          const double* bias_rel, int Q, int is_regressor, const double* iso_x,
                          const double* iso_y, int n_iso, double* out);

/* ---------------------------------------------------------------------------------------------
 * Dual path (single GPU; "replicas only").  Replaces _optimize_alpha_gamma
 * (_neo_ls_svm.py:252-323) for rho = 1 and the dual branches of decision_function (:668-671) and
 * predict_std (:473-475).  The n x G x n tensor of :272-282 is never formed.
 * ------------------------------------------------------------------------------------------- */
/* Kernel matrix F = exp(-0.5 ||x_i - x_j||^2) + 1 (:261), lam, Q = eigh(sn F sn) (:265), and the LOO
 * sweep over `gammas` (:268-302).  Xt: n x p transformed training rows; s: weights / sum (:252);
 * sn = s / median|s| (:253).  Outputs: sums_out 3 x G (as nls_primal_loo_sweep), yhat_loo_out n x G
 * (the LOO predictions at every gamma, :286), lam_out n.  F, Q and F0.SQ stay in the context for
 * nls_dual_finalize. */
int nls_dual_sweep(nls_ctx* ctx, const double* Xt, int n, int p, const double* y, const double* s,
                   const double* sn, const double* gammas, int G, int is_classifier,
                   double* sums_out, double* yhat_loo_out, double* lam_out);
/* Per-row outputs at the selected gamma (:311-323); must follow nls_dual_sweep with the same n.
 * alpha_out: Cholesky re-solve (:313-314) when U_out != NULL (U_out: n x n, cho_factor layout of
 * gamma diag(sn^-2) + F), otherwise the eigen-expansion; alpha_eig_out (optional): eigen-expansion
 * (:311); Falpha_out (optional): F alpha (:315); sigma2_out (optional): 1 - k_i^T (.)^-1 k_i (:322);
 * Bt_out (optional, n x n) and w_out (optional, n): (gamma S^-2 + F)^-1 = Bt^T diag(w) Bt for
 * nls_dual_predict. */
int nls_dual_finalize(nls_ctx* ctx, int n, const double* y, const double* sn, double gamma,
                      double* alpha_out, double* alpha_eig_out, double* U_out, double* Falpha_out,
                      double* sigma2_out, double* Bt_out, double* w_out);
/* yhat = K(xq, Xt) alpha + alpha_sum (:669-671); sigma = sqrt(1 - sum_k ((K Bt^T)_ik)^2 w_k) (:474-477).
 * Either output may be NULL. */
int nls_dual_predict(nls_ctx* ctx, const double* Xq, int64_t nq, const double* Xt, int n, int p,
                     const double* alpha, double alpha_sum, const double* Bt, const double* w,
                     double* yhat_out, double* sigma_out);

/* ---------------------------------------------------------------------------------------------
 * Supervised affine pre-pass (SURVEY.md §8f "next" #1): per-bin weighted median and weighted mean
 * absolute deviation of every feature.  Replaces the host argsort of every column of every target
 * bin in AffineNormalizer.fit (_affine_normalizer.py:81-88, _weighted_quantile.py:52-63).
 *   perm (n int64): row indices grouped by bin; w (n): per-row weight, normalised within its bin, in
 *   perm order; tiles (ntiles x 4 int32: bin, first, last+1, 0): row ranges of perm that do not
 *   straddle bins; bin_tiles (nbins x 2 int32): tile range of every bin.
 * nls_bin_median_stats writes stats_out[7][nbins][d] around the value v* at which the cumulative
 * weight first exceeds half the bin's weight: v*, predecessor value (NaN if none), successor value
 * (NaN if none), weight strictly below v*, weight of the ties at v*, number of ties, weight of the
 * first tie; and wtot_out[nbins][d] = total weight.  The host evaluates the reference's two linear
 * interpolations from these.  thresh (nbins, may be NULL): crossing threshold of every bin in place of half its
 * total weight; with unit weights and thresh[b] = r + 0.5, v* is the order statistic of rank r (0-based), which is
 * how the host reproduces the reference's interpolation bit for bit for uniformly weighted bins.
 * nls_bin_mad: spread_out[bin][col] = sum_i w_i |x_i,col - centre[bin][col]|.
 * ------------------------------------------------------------------------------------------- */
int nls_bin_median_stats(nls_ctx* ctx, const double* X, int64_t n, int d, const int64_t* perm,
                         const double* w, const int* tiles, int ntiles, const int* bin_tiles,
                         int nbins, const double* thresh, double* stats_out, double* wtot_out);
int nls_bin_mad(nls_ctx* ctx, const double* X, int64_t n, int d, const int64_t* perm, const double* w,
                const int* tiles, int ntiles, const int* bin_tiles, int nbins, const double* centre,
                double* spread_out);

/* Micro-benchmarks used for the roofline denominators (bench.py / profiles/). */
int nls_bench_dmma_peak(nls_ctx* ctx, int iters, double* tflops_out);
/* INT8 tensor rate (TOP/s) of a resident-tile tcgen05.mma kind::i8 loop on every SM: n_cols = 256 (the hardware
 * peak) or 64 (the shape of the Ozaki tile, bound by the shared-memory operand reads of each MMA). */
int nls_bench_i8_peak(nls_ctx* ctx, int iters, int n_cols, double* tops_out);

#ifdef __cplusplus
}
#endif
#endif /* NLS_B200_H */
