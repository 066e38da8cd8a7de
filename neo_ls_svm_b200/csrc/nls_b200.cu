// nls_b200.cu — host side of libnls_b200.so: context, scratch management, stage drivers and the
// C ABI declared in include/nls_b200.h.  sm_100a only; there is no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nls_b200.h"
#include "ops.cuh"
#include "small_kernels.cuh"
#include "jacobi.cuh"
#include "jacobi_wide.cuh"
#include "binstats.cuh"
#include "stedc_host.h"
#include "stedc.cuh"
#include "hetrd.cuh"
#include "potrf.cuh"
#include "ozaki.cuh"

using namespace nls;

// ---------------------------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(NLS_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

#define SOLVER_TRY(expr)                                                                          \
  do {                                                                                            \
    cusolverStatus_t _s = (expr);                                                                 \
    if (_s != CUSOLVER_STATUS_SUCCESS)                                                            \
      return fail(NLS_ERR_SOLVER, "%s failed at %s:%d: status %d", #expr, __FILE__, __LINE__, (int)_s); \
  } while (0)

#define NLS_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != NLS_OK) return _r; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// Scratch of the tridiagonalisation / divide-and-conquer eigensolver (csrc/eig_driver.inc), grow-only.
struct EigBuffers {
  DevBuf aw, vw, pwrw, vtvr, vecs, tfac, part, y, qa, qb, qp, u1, u2, delta, dc_d, dc_i, dc_desc, dc_rot, ws, c2, chol;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct ProfSpan {
  cudaEvent_t a, b;
  int kind;
};

struct nls_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t chunk_rows = 32768;
  int64_t launches = 0;
  bool use_tma = true;
  int eig_kind = 2;  // 0: hand-written block Jacobi, 1: cuSOLVER Zheevd (library comparator), 2: auto
  int eig_sweeps = 0;  // sweeps used by the last Jacobi solve
  cudaStream_t jac_stream = nullptr;  // side stream (graph capture is not allowed on the legacy default stream)
  cudaGraphExec_t jac_graph = nullptr;
  const void* jac_graph_key = nullptr;   // the captured graph bakes in the pointers of both Jacobi buffers:
  const void* jac_graph_key2 = nullptr;  // jac_mat (key) and jac_small (key2); either one moving invalidates it
  int jac_graph_nb = 0;
  int jac_inner = 1;  // cyclic sweeps per 8x8 pivot solve (partial diagonalisation is enough for block Jacobi)
  bool jac_wide_attr = false;  // dynamic shared memory opt-in of the wide Jacobi kernels done on this device
  cudaStream_t copy_stream = nullptr;        // host -> device row groups of nls_primal_gram_h2d
  std::vector<cudaEvent_t> copy_events;      // one per row group, reused
  EncodeTiledFn encode = nullptr;
  cusolverDnHandle_t solver = nullptr;
  // scratch (grow-only, zero-filled when (re)allocated)
  DevBuf xc, wt, psi, psiT, pu, bt, rt, small, part, gram_ws, border, rowtmp, solver_ws, solver_mat;
  // dual path (state kept between nls_dual_sweep and nls_dual_finalize)
  DevBuf jac_mat, jac_small, bs_part, bs_keys, dotpart;
  DevBuf d_xpad, d_xq, d_norm, d_fm, d_sq, d_sqt, d_g1, d_ab, d_ra, d_vec, d_ng, d_kq, d_btp;
  // INT8 (Ozaki) GEMM core of the projection (csrc/ozaki.cuh): digit planes of the chunk / of the basis, per-column
  // exponents and recombination scales
  DevBuf oz_a, oz_b, oz_g, oz_small, oz_sa, oz_sb, oz_ssmall;
  int gemm_core = 2;      // 2 (default): Gram, projection, predict_std AND the γ sweep on tcgen05 kind::i8 (Ozaki scheme),
                          // 1: the sweep stays on DMMA, 0: everything on DMMA
  bool oz_attr = false;   // dynamic shared memory opt-in of the INT8 kernels done on this device
  int dual_n = 0;
  const void* dual_y = nullptr;   // the y / sn the pending dual sweep was run with: nls_dual_finalize must be given
  const void* dual_sn = nullptr;  // the same ones (guards against interleaving two models on one context)
  EigBuffers eig;
  std::vector<double> last_d, last_e;  // tridiagonal form produced by the last eigensolve (host copies)
  // profiling
  bool prof = false;
  std::vector<ProfSpan> spans;
  double prof_ms[NLS_PROF_N] = {0};
  int64_t prof_n[NLS_PROF_N] = {0};
};

static int ensure(nls_ctx* ctx, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return NLS_OK;
  if (b.p) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.bytes = 0;
  }
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) return fail(NLS_ERR_ALLOC, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  b.bytes = bytes;
  CUDA_TRY(cudaMemsetAsync(b.p, 0, bytes, ctx->stream));
  return NLS_OK;
}

static int oz_attr(nls_ctx* ctx);  // dynamic shared memory opt-in of the INT8 GEMM kernels (defined with the projection)

static inline long long round_up(long long x, long long q) { return (x + q - 1) / q * q; }
static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  return (int)std::max<long long>(1, std::min<long long>(g, cap));
}

struct ProfScope {
  nls_ctx* ctx;
  int kind;
  ProfSpan span;
  bool on;
  ProfScope(nls_ctx* c, int k) : ctx(c), kind(k), on(c->prof) {
    if (on) {
      cudaEventCreate(&span.a);
      cudaEventCreate(&span.b);
      span.kind = kind;
      cudaEventRecord(span.a, ctx->stream);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(span.b, ctx->stream);
      ctx->spans.push_back(span);
    }
  }
};

static int check_launch(nls_ctx* ctx, const char* what) {
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(NLS_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return NLS_OK;
}

// ---------------------------------------------------------------------------------------------
// Tensor maps and GEMM launch
// ---------------------------------------------------------------------------------------------
static int make_map(nls_ctx* ctx, CUtensorMap* map, const Operand& op, int box_rows, long long total_rows) {
  if (!ctx->use_tma) {
    memset(map, 0, sizeof *map);
    return NLS_OK;
  }
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) || (op.ld & 1))
    return fail(NLS_ERR_INVALID, "TMA operand must be 16-byte aligned with an even pitch");
  cuuint64_t dims[2] = {(cuuint64_t)op.kext, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)op.ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(op.ptr), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(NLS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return NLS_OK;
}

// `rows_a_total` / `rows_b_total`: row extent of the whole 2-D array each map describes (all planes).
// For MODE_DUAL_A/COMPLEX the K extent of the map covers both planes when plane_dk != 0.
template <int MODE, class Op>
static int launch_gemm(nls_ctx* ctx, const typename Op::Params& p, dim3 grid, long long rows_a_total,
                       long long rows_b_total, int kind, const char* name) {
  using T = ModeTraits<MODE>;
  Operand a = p.A, b = p.B;
  // The tensor map's inner extent must span plane b when it lives at a K offset.
  Operand am = a, bm = b;
  if (T::A_PLANES == 2 && a.plane_dk) am.kext = a.plane_dk + a.kext;
  if (T::B_PLANES == 2 && b.plane_dk) bm.kext = b.plane_dk + b.kext;
  CUtensorMap mapA, mapB;
  NLS_TRY(make_map(ctx, &mapA, am, BM, rows_a_total));
  NLS_TRY(make_map(ctx, &mapB, bm, BN, rows_b_total));
  ProfScope scope(ctx, kind);
  if (ctx->use_tma) {
    auto kern = gemm_kernel<MODE, Op, true>;
    static bool attr_set[64] = {false};  // function attributes are per device
    if (!attr_set[ctx->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
      attr_set[ctx->device & 63] = true;
    }
    kern<<<grid, GEMM_THREADS, T::SMEM_BYTES, ctx->stream>>>(mapA, mapB, p);
  } else {
    auto kern = gemm_kernel<MODE, Op, false>;
    static bool attr_set[64] = {false};  // function attributes are per device
    if (!attr_set[ctx->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES));
      attr_set[ctx->device & 63] = true;
    }
    kern<<<grid, GEMM_THREADS, T::SMEM_BYTES, ctx->stream>>>(mapA, mapB, p);
  }
  return check_launch(ctx, name);
}

// ---------------------------------------------------------------------------------------------
// Stage drivers (internal)
// ---------------------------------------------------------------------------------------------
struct MapGeom {
  int d, D, dpad;
  int Dp;   // K padding of the planar chunk (multiple of 16)
  int DpT;  // row padding of each plane of the transposed chunk (multiple of 128)
  int m, Np;
  long long ldp;  // pitch of P/U/r^T rows (multiple of 16)
};

static MapGeom geom(int d, int D) {
  MapGeom g;
  g.d = d;
  g.D = D;
  g.dpad = (int)round_up(d, 16);
  g.Dp = (int)round_up(D, 16);
  g.DpT = (int)round_up(D, 128);
  g.m = D + 1;
  g.Np = (int)round_up(D + 1, 64);
  g.ldp = round_up(D + 1, 16);
  return g;
}

// W (d x D) -> W^T (D x dpad) scratch.
static int prep_weights(nls_ctx* ctx, const MapGeom& g, const double* W) {
  NLS_TRY(ensure(ctx, ctx->wt, (size_t)g.D * g.dpad * 8));
  NLS_TRY(ensure(ctx, ctx->xc, (size_t)ctx->chunk_rows * g.dpad * 8));
  transpose_w_kernel<<<grid_for((long long)g.D * g.dpad), 256, 0, ctx->stream>>>(W, g.d, g.D, g.dpad,
                                                                                (double*)ctx->wt.p);
  return check_launch(ctx, "transpose_w_kernel");
}

// Feature map of `rows` rows starting at X into `out` with the given layout.
static int feature_chunk(nls_ctx* ctx, const MapGeom& g, const double* X, const double* shift, int rows, int layout,
                         double* out, long long ld, int plane_stride, const double* row_scale,
                         const RowDots* dots = nullptr) {
  center_rows_kernel<<<grid_for((long long)rows * g.dpad), 256, 0, ctx->stream>>>(X, shift, rows, g.d, g.dpad,
                                                                                 (double*)ctx->xc.p);
  NLS_TRY(check_launch(ctx, "center_rows_kernel"));
  OpFeatureMap::Params p;
  p.A = Operand{(const double*)ctx->xc.p, g.dpad, rows, g.d, 0, 0};
  p.B = Operand{(const double*)ctx->wt.p, g.dpad, g.D, g.d, 0, 0};
  p.n_rows = rows;
  p.D = g.D;
  p.layout = layout;
  p.inv_sqrt_D = 1.0 / sqrt((double)g.D);
  p.row_scale = row_scale;
  p.out = out;
  p.ld = ld;
  p.plane_stride = plane_stride;
  if (dots) {
    p.dots = *dots;
  } else {
    p.dots = RowDots{};
    p.dots.count = 0;
  }
  dim3 grid((g.D + BN - 1) / BN, (rows + BM - 1) / BM);
  return launch_gemm<MODE_REAL, OpFeatureMap>(ctx, p, grid, rows, g.D, NLS_PROF_FEATURE_MAP, "feature_map");
}

// Partials buffer of the fused row dot products: [feature tiles][2][chunk rows].
static int rowdots_buffer(nls_ctx* ctx, const MapGeom& g, double** out) {
  const size_t tiles = (size_t)(g.D + BN - 1) / BN;
  NLS_TRY(ensure(ctx, ctx->dotpart, tiles * 2 * (size_t)ctx->chunk_rows * 8));
  *out = (double*)ctx->dotpart.p;
  return NLS_OK;
}

// Number of leading projection columns computed by 64-wide GEMM tiles.  When only a few columns spill into
// a last tile (m = D + 1 with D a multiple of 64 leaves exactly one), they go to project_tail_kernel.
static inline int tail_split(int m) { return (m % BN != 0 && m % BN <= 4 && m > BN) ? m - m % BN : m; }

// Basis planes B^T (2Np x Dp) + constant-feature bias from a numpy-layout m x m complex basis.
struct BasisScratch {
  double *bt, *bias_r, *bias_i, *v_r, *v_i, *w;
};
static int prep_basis(nls_ctx* ctx, const MapGeom& g, const double* B, BasisScratch* out, int upper = 0) {
  const size_t bt_bytes = (size_t)2 * g.Np * g.Dp * 8;
  NLS_TRY(ensure(ctx, ctx->bt, bt_bytes));
  NLS_TRY(ensure(ctx, ctx->small, (size_t)8 * g.Np * 8));
  double* sm = (double*)ctx->small.p;
  out->bt = (double*)ctx->bt.p;
  out->bias_r = sm;
  out->bias_i = sm + g.Np;
  out->v_r = sm + 2 * g.Np;
  out->v_i = sm + 3 * g.Np;
  out->w = sm + 4 * g.Np;
  dim3 grid((g.m + 31) / 32, (g.m + 31) / 32), block(32, 8);
  split_basis_kernel<<<grid, block, 0, ctx->stream>>>(B, g.m, g.D, g.Np, g.Dp, upper, out->bt, out->bias_r, out->bias_i);
  return check_launch(ctx, "split_basis_kernel");
}

static Operand psi_operand(nls_ctx* ctx, const MapGeom& g, int rows) {
  return Operand{(const double*)ctx->psi.p, 2LL * g.Dp, rows, g.D, 0, g.Dp};
}
static Operand basis_operand(const MapGeom& g, const double* bt) {
  return Operand{bt, g.Dp, 2 * g.Np, g.D, g.Np, 0};
}

// ---------------------------------------------------------------------------------------------
// C ABI: context
// ---------------------------------------------------------------------------------------------
extern "C" int nls_version(void) { return 100; }
extern "C" const char* nls_last_error(void) { return g_last_error.c_str(); }

extern "C" int nls_ctx_create(int device, void* stream, nls_ctx** out) {
  if (!out) return fail(NLS_ERR_INVALID, "out is null");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(NLS_ERR_ARCH, "device %d is sm_%d%d; this library only contains sm_100a code", device, prop.major,
                prop.minor);
  nls_ctx* ctx = new nls_ctx();
  ctx->device = device;
  ctx->stream = (cudaStream_t)stream;
  ctx->sm_count = prop.multiProcessorCount;
  const char* env = getenv("NLS_NO_TMA");
  ctx->use_tma = !(env && env[0] == '1');
  env = getenv("NLS_EIG");
  if (env && strcmp(env, "cusolver") == 0) ctx->eig_kind = 1;
  if (env && strcmp(env, "jacobi") == 0) ctx->eig_kind = 0;
  if (env && strcmp(env, "dc") == 0) ctx->eig_kind = 3;
  env = getenv("NLS_GEMM");
  if (env && strcmp(env, "dmma") == 0) ctx->gemm_core = 0;
  if (env && strcmp(env, "ozaki") == 0) ctx->gemm_core = 2;
  if (env && strcmp(env, "ozaki-dmma-sweep") == 0) ctx->gemm_core = 1;
  env = getenv("NLS_JACOBI_INNER");
  if (env && atoi(env) > 0) ctx->jac_inner = atoi(env);
  env = getenv("NLS_CHUNK_ROWS");
  if (env && atoll(env) > 0) ctx->chunk_rows = round_up(atoll(env), 128);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete ctx;
    return fail(NLS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  }
  ctx->encode = (EncodeTiledFn)fn;
  if (cusolverDnCreate(&ctx->solver) != CUSOLVER_STATUS_SUCCESS) {
    delete ctx;
    return fail(NLS_ERR_SOLVER, "cusolverDnCreate failed");
  }
  cusolverDnSetStream(ctx->solver, ctx->stream);
  *out = ctx;
  return NLS_OK;
}

extern "C" int nls_ctx_destroy(nls_ctx* ctx) {
  if (!ctx) return NLS_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->xc, &ctx->wt, &ctx->psi, &ctx->psiT, &ctx->pu, &ctx->bt, &ctx->rt, &ctx->small,
                    &ctx->part, &ctx->gram_ws, &ctx->border, &ctx->rowtmp, &ctx->solver_ws, &ctx->solver_mat,
                    &ctx->jac_mat, &ctx->jac_small, &ctx->bs_part, &ctx->bs_keys, &ctx->dotpart, &ctx->d_xpad, &ctx->d_xq, &ctx->d_norm, &ctx->d_fm, &ctx->d_sq, &ctx->d_sqt, &ctx->d_g1,
                    &ctx->d_ab, &ctx->d_ra, &ctx->d_vec, &ctx->d_ng, &ctx->d_kq, &ctx->d_btp, &ctx->oz_a, &ctx->oz_b, &ctx->oz_g, &ctx->oz_small, &ctx->oz_sa, &ctx->oz_sb, &ctx->oz_ssmall};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  {
    EigBuffers& e = ctx->eig;
    DevBuf* eb[] = {&e.aw, &e.vw, &e.pwrw, &e.vtvr, &e.vecs, &e.tfac, &e.part, &e.y, &e.qa, &e.qb, &e.qp, &e.u1, &e.u2,
                    &e.delta, &e.dc_d, &e.dc_i, &e.dc_desc, &e.dc_rot, &e.ws, &e.c2, &e.chol};
    for (DevBuf* b : eb)
      if (b->p) cudaFree(b->p);
  }
  for (auto& s : ctx->spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  if (ctx->solver) cusolverDnDestroy(ctx->solver);
  if (ctx->jac_graph) cudaGraphExecDestroy(ctx->jac_graph);
  if (ctx->jac_stream) cudaStreamDestroy(ctx->jac_stream);
  for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return NLS_OK;
}

extern "C" int nls_ctx_set_stream(nls_ctx* ctx, void* stream) {
  if (!ctx) return fail(NLS_ERR_INVALID, "ctx is null");
  CUDA_TRY(cudaSetDevice(ctx->device));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // scratch buffers are shared: drain the old stream first
  ctx->stream = (cudaStream_t)stream;
  if (ctx->solver) cusolverDnSetStream(ctx->solver, ctx->stream);
  return NLS_OK;
}

extern "C" int nls_ctx_set_chunk_rows(nls_ctx* ctx, int64_t rows) {
  if (!ctx || rows < 128) return fail(NLS_ERR_INVALID, "chunk rows must be >= 128");
  ctx->chunk_rows = round_up(rows, 128);
  return NLS_OK;
}

extern "C" int64_t nls_ctx_launch_count(const nls_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int nls_ctx_profile(nls_ctx* ctx, int enable) {
  if (!ctx) return fail(NLS_ERR_INVALID, "ctx is null");
  ctx->prof = enable != 0;
  for (auto& s : ctx->spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  ctx->spans.clear();
  for (int k = 0; k < NLS_PROF_N; ++k) {
    ctx->prof_ms[k] = 0;
    ctx->prof_n[k] = 0;
  }
  return NLS_OK;
}

extern "C" int nls_ctx_profile_read(nls_ctx* ctx, double* ms_out, int64_t* launches_out) {
  if (!ctx) return fail(NLS_ERR_INVALID, "ctx is null");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (auto& s : ctx->spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.a, s.b);
    ctx->prof_ms[s.kind] += ms;
    ctx->prof_n[s.kind] += 1;
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  ctx->spans.clear();
  for (int k = 0; k < NLS_PROF_N; ++k) {
    if (ms_out) ms_out[k] = ctx->prof_ms[k];
    if (launches_out) launches_out[k] = ctx->prof_n[k];
  }
  return NLS_OK;
}

static int check_map_args(nls_ctx* ctx, const void* X, int64_t n, int d, const void* shift, const void* W, int D) {
  if (!ctx) return fail(NLS_ERR_INVALID, "ctx is null");
  if (!X || !shift || !W) return fail(NLS_ERR_INVALID, "null input pointer");
  if (n < 1 || d < 1 || D < 1) return fail(NLS_ERR_INVALID, "n, d and D must be positive (n=%lld d=%d D=%d)", (long long)n, d, D);
  if (n > (1LL << 40)) return fail(NLS_ERR_INVALID, "n too large");
  CUDA_TRY(cudaSetDevice(ctx->device));
  return NLS_OK;
}

// ---------------------------------------------------------------------------------------------
// Stage 1
// ---------------------------------------------------------------------------------------------
extern "C" int nls_feature_map(nls_ctx* ctx, const double* X, int64_t n, int d, const double* shift, const double* W,
                               int D, double* phi_out) {
  NLS_TRY(check_map_args(ctx, X, n, d, shift, W, D));
  if (!phi_out) return fail(NLS_ERR_INVALID, "phi_out is null");
  const MapGeom g = geom(d, D);
  NLS_TRY(prep_weights(ctx, g, W));
  for (int64_t i0 = 0; i0 < n; i0 += ctx->chunk_rows) {
    const int rows = (int)std::min<int64_t>(ctx->chunk_rows, n - i0);
    NLS_TRY(feature_chunk(ctx, g, X + i0 * d, shift, rows, FM_COMPLEX, phi_out + i0 * (D + 1) * 2, 0, 0, nullptr));
  }
  return NLS_OK;
}

// z = (x - shift) W as a plain real matrix (the dual path's feature map and AffineFeatureMap.transform).
extern "C" int nls_affine_map(nls_ctx* ctx, const double* X, int64_t n, int d, const double* shift, const double* W,
                              int D, double* Z_out) {
  NLS_TRY(check_map_args(ctx, X, n, d, shift, W, D));
  if (!Z_out) return fail(NLS_ERR_INVALID, "Z_out is null");
  const MapGeom g = geom(d, D);
  NLS_TRY(prep_weights(ctx, g, W));
  for (int64_t i0 = 0; i0 < n; i0 += ctx->chunk_rows) {
    const int rows = (int)std::min<int64_t>(ctx->chunk_rows, n - i0);
    center_rows_kernel<<<grid_for((long long)rows * g.dpad), 256, 0, ctx->stream>>>(X + i0 * d, shift, rows, g.d, g.dpad,
                                                                                   (double*)ctx->xc.p);
    NLS_TRY(check_launch(ctx, "center_rows_kernel"));
    OpStore<false>::Params p;
    p.A = Operand{(const double*)ctx->xc.p, g.dpad, rows, g.d, 0, 0};
    p.B = Operand{(const double*)ctx->wt.p, g.dpad, g.D, g.d, 0, 0};
    p.n_rows = rows;
    p.n_cols = D;
    p.norm_a = nullptr;
    p.norm_b = nullptr;
    p.add = 0.0;
    p.zero_diag_to = -1;
    p.out = Z_out + i0 * D;
    p.ld = D;
    NLS_TRY((launch_gemm<MODE_REAL, OpStore<false>>(ctx, p, dim3((D + BN - 1) / BN, (rows + BM - 1) / BM), rows, D,
                                                     NLS_PROF_FEATURE_MAP, "affine_map")));
  }
  return NLS_OK;
}

// ---------------------------------------------------------------------------------------------
// Stage 2
// ---------------------------------------------------------------------------------------------
// Host rows streamed into the device buffers while the Gram pass runs (nls_primal_gram_h2d).
struct UploadPlan {
  const double *Xh, *yh, *sh;
  double *Xd, *yd, *sd;
  int64_t group_rows;
};

static int gram_impl(nls_ctx* ctx, const double* X, const double* y, const double* s, int64_t n, int d,
                     const double* shift, const double* W, int D, double* A_out, double* b_out,
                     const UploadPlan* up) {
  NLS_TRY(check_map_args(ctx, X, n, d, shift, W, D));
  if (!y || !s || !A_out || !b_out) return fail(NLS_ERR_INVALID, "null pointer");
  if (up) {
    // Enqueue every row group on the copy stream now (the copy engine works through them in order) with one
    // event per group; the compute stream waits for a group's event right before its first chunk.
    if (!ctx->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    const int64_t groups = (n + up->group_rows - 1) / up->group_rows;
    while ((int64_t)ctx->copy_events.size() < groups) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->copy_events.push_back(e);
    }
    // The destination buffers may still be in use by work queued on the compute stream (allocator reuse).
    cudaEvent_t fence = ctx->copy_events[0];
    CUDA_TRY(cudaEventRecord(fence, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, fence, 0));
    for (int64_t gi = 0; gi < groups; ++gi) {
      const int64_t r0 = gi * up->group_rows, rows = std::min<int64_t>(up->group_rows, n - r0);
      CUDA_TRY(cudaMemcpyAsync(up->Xd + r0 * d, up->Xh + r0 * d, (size_t)rows * d * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      CUDA_TRY(cudaMemcpyAsync(up->yd + r0, up->yh + r0, (size_t)rows * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      CUDA_TRY(cudaMemcpyAsync(up->sd + r0, up->sh + r0, (size_t)rows * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      CUDA_TRY(cudaEventRecord(ctx->copy_events[gi], ctx->copy_stream));
    }
  }
  const MapGeom g = geom(d, D);
  NLS_TRY(prep_weights(ctx, g, W));
  const long long ldT = ctx->chunk_rows;
  NLS_TRY(ensure(ctx, ctx->psiT, (size_t)2 * g.DpT * ldT * 8));
  const int tiles_m = (D + BM - 1) / BM, tiles_n = (D + BN - 1) / BN;
  int n_tiles = 0;
  for (int kb = 0; kb < tiles_m; ++kb) n_tiles += std::max(0, tiles_n - 2 * kb);
  // INT8 (Ozaki) core: 128-feature row blocks x 32-feature column tiles, upper triangle only; the chunk's rows are K.
  const bool use_oz = ctx->gemm_core >= 1;
  const int oz_rb = (D + oz::TM - 1) / oz::TM, oz_nt = (D + oz::TN / 2 - 1) / (oz::TN / 2);
  int oz_tiles = 0;
  for (int kb = 0; kb < oz_rb; ++kb) oz_tiles += oz_nt - 4 * kb;
  // K (= 2 x the chunk's rows) is split so that one work item stays inside the INT32 accumulators (oz::MAX_K) and,
  // for small D, so that every SM has a work item.
  const int oz_min_splits = (int)((2 * round_up(ctx->chunk_rows, oz::KS) + oz::MAX_K - 1) / oz::MAX_K);
  const int splits = use_oz ? std::max(oz_min_splits, ctx->sm_count / oz_tiles) : std::max(1, ctx->sm_count / n_tiles);
  if (use_oz) {
    NLS_TRY(oz_attr(ctx));
    const size_t nks2_max = (size_t)2 * (round_up(ctx->chunk_rows, oz::KS) / oz::KS);
    NLS_TRY(ensure(ctx, ctx->oz_a, (size_t)oz_rb * nks2_max * oz::A_STAGE));
    NLS_TRY(ensure(ctx, ctx->oz_g, (size_t)oz_nt * nks2_max * oz::B_STAGE));
    NLS_TRY(ensure(ctx, ctx->oz_small, 64));
  }
  const size_t ws_bytes = (size_t)splits * 2 * D * D * 8;
  NLS_TRY(ensure(ctx, ctx->gram_ws, ws_bytes));
  NLS_TRY(ensure(ctx, ctx->border, (size_t)(4 * D + 2) * 8));
  CUDA_TRY(cudaMemsetAsync(ctx->gram_ws.p, 0, ws_bytes, ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ctx->border.p, 0, (size_t)(4 * D + 2) * 8, ctx->stream));
  double* border = (double*)ctx->border.p;
  double* scal = border + 4 * D;
  int64_t arrived = 0;  // rows whose upload the compute stream has been ordered after
  for (int64_t i0 = 0; i0 < n; i0 += ctx->chunk_rows) {
    const int rows = (int)std::min<int64_t>(ctx->chunk_rows, n - i0);
    while (up && arrived < i0 + rows) {
      CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[arrived / up->group_rows], 0));
      arrived = std::min<int64_t>(n, (arrived / up->group_rows + 1) * up->group_rows);
    }
    NLS_TRY(feature_chunk(ctx, g, X + i0 * d, shift, rows, FM_TRANSPOSED, (double*)ctx->psiT.p, ldT, g.DpT, s + i0));
    if (use_oz) {
      double* sc = (double*)ctx->oz_small.p;
      const int nksR = (int)(round_up(rows, oz::KS) / oz::KS);
      {
      ProfScope slice_scope(ctx, NLS_PROF_SLICE);
      oz::gram_scale_kernel<<<1, 1024, 0, ctx->stream>>>(s + i0, rows, 1.0 / sqrt((double)D), sc);
      NLS_TRY(check_launch(ctx, "oz::gram_scale_kernel"));
      oz::slice_gram_kernel<oz::IMAGE><<<grid_for((long long)oz_rb * 2 * nksR * oz::TM * 2), 256, 0, ctx->stream>>>(
          (const double*)ctx->psiT.p, ldT, g.DpT, D, rows, sc, nksR, oz_rb, oz_nt, (int8_t*)ctx->oz_a.p, (int8_t*)ctx->oz_g.p);
      NLS_TRY(check_launch(ctx, "oz::slice_gram_kernel"));
      }
      ProfScope scope(ctx, NLS_PROF_GRAM);
      const int ks_per_split = (2 * nksR + splits - 1) / splits;
      oz::GemmParams gp{(const int8_t*)ctx->oz_a.p, (const int8_t*)ctx->oz_g.p, 2 * nksR, oz_rb, oz_nt, 1, oz_tiles,
                        (2 * nksR + ks_per_split - 1) / ks_per_split, ks_per_split, 0, 1, 1 << 30, 0};
      oz::EpiGram::Params ep{D, sc + 1, (double*)ctx->gram_ws.p};
      const int grid = (int)std::min<long long>((long long)gp.tiles * gp.splits, ctx->sm_count);
      oz::gemm_kernel_i8<oz::IMAGE, oz::EpiGram><<<grid, oz::THREADS, oz::SMEM_BYTES, ctx->stream>>>(gp, ep);
      NLS_TRY(check_launch(ctx, "oz::gemm_kernel_i8<EpiGram>"));
    } else {
    OpGram::Params p;
    p.A = Operand{(const double*)ctx->psiT.p, ldT, 2 * g.DpT, rows, g.DpT, 0};
    p.B = p.A;
    p.D = D;
    p.n_tiles_m = tiles_m;
    p.n_tiles_n = tiles_n;
    p.rows = rows;
    p.k_per_split = (int)round_up((rows + splits - 1) / splits, BK);
    p.ws = (double*)ctx->gram_ws.p;
    NLS_TRY((launch_gemm<MODE_COMPLEX, OpGram>(ctx, p, dim3(n_tiles, splits), 2 * g.DpT, 2 * g.DpT, NLS_PROF_GRAM,
                                                "gram")));
    }
    gram_border_kernel<<<2 * D + 1, 256, 0, ctx->stream>>>((const double*)ctx->psiT.p, ldT, D, g.DpT, rows, s + i0,
                                                           y + i0, border, scal);
    NLS_TRY(check_launch(ctx, "gram_border_kernel"));
  }
  gram_assemble_kernel<<<grid_for((long long)g.m * g.m), 256, 0, ctx->stream>>>((const double*)ctx->gram_ws.p, splits,
                                                                               D, border, scal, A_out, b_out);
  return check_launch(ctx, "gram_assemble_kernel");
}

extern "C" int nls_primal_gram(nls_ctx* ctx, const double* X, const double* y, const double* s, int64_t n, int d,
                               const double* shift, const double* W, int D, double* A_out, double* b_out) {
  return gram_impl(ctx, X, y, s, n, d, shift, W, D, A_out, b_out, nullptr);
}

extern "C" int nls_primal_gram_h2d(nls_ctx* ctx, const double* X_host, const double* y_host, const double* s_host,
                                   int64_t n, int d, double* X_dev, double* y_dev, double* s_dev, const double* shift,
                                   const double* W, int D, double* A_out, double* b_out) {
  if (!ctx) return fail(NLS_ERR_INVALID, "ctx is null");
  if (!X_host || !y_host || !s_host || !X_dev || !y_dev || !s_dev) return fail(NLS_ERR_INVALID, "null pointer");
  UploadPlan up{X_host, y_host, s_host, X_dev, y_dev, s_dev, ctx->chunk_rows * 4};
  return gram_impl(ctx, X_dev, y_dev, s_dev, n, d, shift, W, D, A_out, b_out, &up);
}

// ---------------------------------------------------------------------------------------------
// Stage 3: hand-written parallel block Jacobi (default) or cuSOLVER Zheevd (library comparator,
// NLS_EIG=cusolver / nls_ctx_set_eigensolver) behind the same entry point.
// ---------------------------------------------------------------------------------------------
static int heev_cusolver(nls_ctx* ctx, const double* A, int m, double scale, double* lam_out, double* Q_out) {
  if (!ctx || !A || !lam_out || !Q_out || m < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_heev");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const long long mm = (long long)m * m;
  NLS_TRY(ensure(ctx, ctx->solver_mat, (size_t)mm * 16 + 64));
  double* work_mat = (double*)ctx->solver_mat.p;
  int* info = (int*)(work_mat + 2 * mm);
  scale_conj_kernel<<<grid_for(mm), 256, 0, ctx->stream>>>(A, mm, scale, work_mat);
  NLS_TRY(check_launch(ctx, "scale_conj_kernel"));
  int lwork = 0;
  SOLVER_TRY(cusolverDnZheevd_bufferSize(ctx->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m,
                                         (cuDoubleComplex*)work_mat, m, lam_out, &lwork));
  NLS_TRY(ensure(ctx, ctx->solver_ws, (size_t)lwork * 16));
  ProfScope scope(ctx, NLS_PROF_OTHER);
  SOLVER_TRY(cusolverDnZheevd(ctx->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, m,
                              (cuDoubleComplex*)work_mat, m, lam_out, (cuDoubleComplex*)ctx->solver_ws.p, lwork,
                              info));
  // Column-major eigenvectors V[l, k] at work_mat[k*m + l]  ->  row-major Q[l, k].
  dim3 grid((m + 31) / 32, (m + 31) / 32), block(32, 8);
  transpose_square_kernel<2><<<grid, block, 0, ctx->stream>>>(work_mat, m, Q_out);
  NLS_TRY(check_launch(ctx, "transpose_square_kernel"));
  int h_info = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (h_info != 0) return fail(NLS_ERR_SOLVER, "Hermitian eigensolver failed: info = %d", h_info);
  return NLS_OK;
}


// Common tail of the Jacobi drivers: eigenvalues = diag(G), ascending order (stable), Q = V[:, perm].
static int jacobi_finish(nls_ctx* ctx, const double* Gr, const double* Vr, const double* Vi, int mp, int m,
                         double* lam_raw, int* perm, double* lam_out, double* Q_out) {
  jacobi_diag_kernel<<<(mp + 255) / 256, 256, 0, ctx->stream>>>(Gr, mp, mp, lam_raw);
  NLS_TRY(check_launch(ctx, "jacobi_diag_kernel"));
  std::vector<double> h_lam(mp);
  CUDA_TRY(cudaMemcpyAsync(h_lam.data(), lam_raw, (size_t)mp * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  std::vector<int> h_perm(m);
  for (int k = 0; k < m; ++k) h_perm[k] = k;  // pad columns (index >= m) never rotate and are dropped
  std::stable_sort(h_perm.begin(), h_perm.end(), [&](int a, int b) { return h_lam[a] < h_lam[b]; });
  CUDA_TRY(cudaMemcpyAsync(perm, h_perm.data(), (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
  jacobi_gather_kernel<<<grid_for((long long)m * m), 256, 0, ctx->stream>>>(Vr, Vi, mp, m, perm, lam_raw, Q_out, lam_out);
  NLS_TRY(check_launch(ctx, "jacobi_gather_kernel"));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // h_perm must outlive the copy
  return NLS_OK;
}

// Wide-block driver (csrc/jacobi_wide.cuh): 8-wide blocks, ping-pong G, one dependency-free kernel per round; one
// sweep (nb - 1 launches + an off-diagonal census) is replayed as a CUDA graph, and the host reads the rotation
// counter and the census after each sweep, so the iteration stops without a final no-op sweep.
template <int JBW>
static int heev_jacobi_wide(nls_ctx* ctx, const double* A, int m, double scale, double* lam_out, double* Q_out) {
  using C = WideCfg<JBW>;
  CUDA_TRY(cudaSetDevice(ctx->device));
  int nb = (m + JBW - 1) / JBW;
  if (nb & 1) ++nb;
  if (nb < 2) nb = 2;
  const int mp = nb * JBW, np = nb / 2;
  const size_t mm = (size_t)mp * mp;
  const int max_sweeps = 60;
  NLS_TRY(ensure(ctx, ctx->jac_mat, 6 * mm * 8));
  NLS_TRY(ensure(ctx, ctx->jac_small, (size_t)4 * np * C::JSZ * 8 + (size_t)(2 * mp + 8) * 8 +
                                          (size_t)(2 * np + max_sweeps + 16 + mp) * 4));
  double* base = (double*)ctx->jac_mat.p;
  double* Vr = base + 4 * mm;
  double* Vi = base + 5 * mm;
  double* Jbuf = (double*)ctx->jac_small.p;        // [2][np][JSZ]
  double* Sbuf = Jbuf + (size_t)2 * np * C::JSZ;   // [2][np][JSZ]
  double* lam_raw = Sbuf + (size_t)2 * np * C::JSZ;
  double* fro2 = lam_raw + mp;
  double* thr = fro2 + 2;
  int* flags = (int*)(fro2 + 8);        // [2][np]
  int* active = flags + 2 * np;         // [max_sweeps + 2]
  int* misc = active + max_sweeps + 2;  // [0] off-diagonal census, [2] sweeps completed
  int* perm = misc + 8;
  WideArgs wa;
  for (int b = 0; b < 2; ++b)
    for (int c = 0; c < 2; ++c) wa.G[b][c] = base + (size_t)(2 * b + c) * mm;
  wa.Vr = Vr; wa.Vi = Vi;
  wa.ld = mp; wa.nb = nb; wa.max_inner = ctx->jac_inner; wa.sm_count = ctx->sm_count;
  wa.thr = thr; wa.Jbuf = Jbuf; wa.Sbuf = Sbuf; wa.flags = flags; wa.active = active; wa.state = misc + 2;
  ProfScope scope(ctx, NLS_PROF_OTHER);
  jacobi_init_kernel<<<grid_for((long long)mm), 256, 0, ctx->stream>>>(A, m, mp, scale, wa.G[0][0], wa.G[0][1], Vr, Vi);
  NLS_TRY(check_launch(ctx, "jacobi_init_kernel"));
  sumsq_kernel<<<1, 1024, 0, ctx->stream>>>(wa.G[0][0], wa.G[0][1], (long long)mm, fro2);
  NLS_TRY(check_launch(ctx, "sumsq_kernel"));
  jacobi_threshold_kernel<<<1, 1, 0, ctx->stream>>>(fro2, thr);
  CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)(2 * np + max_sweeps + 16) * 4, ctx->stream));
  const size_t piv_smem = (size_t)4 * C::PIV_SM * 8, round_smem = (size_t)wide_smem_doubles<JBW>() * 8;
  if (!ctx->jac_wide_attr) {  // per device (one context per device), not per process
    CUDA_TRY(cudaFuncSetAttribute(jacobi_pivot0_w_kernel<JBW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)piv_smem));
    CUDA_TRY(cudaFuncSetAttribute(jacobi_round_w_kernel<JBW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem));
    ctx->jac_wide_attr = true;
  }
  // Two 8-warp CTAs per SM; warp 0 of the first np CTAs solves a pivot, every other warp updates tiles.
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jacobi_round_w_kernel<JBW>, WIDE_WARPS * 32, round_smem));
  if (occ < 1) return fail(NLS_ERR_CUDA, "jacobi_round_w_kernel does not fit on an SM");
  const long long tasks = (long long)np * (np + 1) / 2 + (long long)(mp / C::P) * np;
  const int grid = (int)std::max<long long>(np, std::min<long long>((long long)ctx->sm_count * occ, (tasks + np + WIDE_WARPS - 1) / WIDE_WARPS));
  if (np > 160) return fail(NLS_ERR_INVALID, "wide Jacobi: m = %d is outside the supported range", m);
  // J(0), S(0): pivots of the very first round.
  jacobi_pivot0_w_kernel<JBW><<<(np + 3) / 4, 128, piv_smem, ctx->stream>>>(wa);
  NLS_TRY(check_launch(ctx, "jacobi_pivot0_w_kernel"));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (!ctx->jac_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->jac_stream, cudaStreamNonBlocking));
  // Timing experiments only (results are then meaningless): bit 0 skips the pivot solves, bit 1 the tile updates.
  // Bit 2 idles warps 4 and 5 of the pivot CTAs (they share the two pivot warps' SM sub-partitions): 24.4 vs 24.6 ms
  // at m = 1025, 7.4 vs 8.0 ms at m = 513, but 190 vs 176 ms at m = 2049 where the tile updates dominate; bit 3
  // additionally idles the presumably co-resident CTA's warps and never paid off (profiles/r1d_jacobi_*.log).
  const char* diag_env = getenv("NLS_JACOBI_DIAG");
  const int diag = diag_env ? atoi(diag_env) & 15 : (np <= 100 ? 4 : 0);
  const int graph_code = -(1 << 24) - ((nb * 64 + JBW * 4) * 16 + diag);
  if (!ctx->jac_graph || ctx->jac_graph_key != (const void*)base || ctx->jac_graph_key2 != (const void*)ctx->jac_small.p || ctx->jac_graph_nb != graph_code) {
    if (ctx->jac_graph) {
      cudaGraphExecDestroy(ctx->jac_graph);
      ctx->jac_graph = nullptr;
    }
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(ctx->jac_stream, cudaStreamCaptureModeRelaxed));
    for (int round = 0; round < nb - 1; ++round)
      jacobi_round_w_kernel<JBW><<<grid, WIDE_WARPS * 32, round_smem, ctx->jac_stream>>>(wa, round, diag);
    cudaMemsetAsync(misc, 0, sizeof(int), ctx->jac_stream);
    jacobi_offdiag_count_w_kernel<<<grid_for((long long)mm), 256, 0, ctx->jac_stream>>>(wa, mp, misc);
    jacobi_sweep_done_w_kernel<<<1, 1, 0, ctx->jac_stream>>>(wa.state);
    CUDA_TRY(cudaStreamEndCapture(ctx->jac_stream, &graph));
    cudaError_t ge = cudaGraphInstantiate(&ctx->jac_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ge != cudaSuccess) return fail(NLS_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ge));
    ctx->jac_graph_key = (const void*)base;
    ctx->jac_graph_key2 = (const void*)ctx->jac_small.p;
    ctx->jac_graph_nb = graph_code;
  }
  int sweep = 0, h_active = 1, h_misc[4] = {1, 0, 0, 0};
  const bool timing_only = (diag & 3) != 0;  // bits 2 and 3 only change the warp layout: results stay valid
  for (; sweep < (timing_only ? 10 : max_sweeps) && (timing_only || (h_active > 0 && h_misc[0] > 0)); ++sweep) {
    CUDA_TRY(cudaGraphLaunch(ctx->jac_graph, ctx->jac_stream));
    ctx->launches += nb;
    CUDA_TRY(cudaMemcpyAsync(&h_active, active + sweep, sizeof(int), cudaMemcpyDeviceToHost, ctx->jac_stream));
    CUDA_TRY(cudaMemcpyAsync(h_misc, misc, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->jac_stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->jac_stream));
  }
  ctx->eig_sweeps = sweep;
  if (h_active > 0 && h_misc[0] > 0 && !timing_only)
    return fail(NLS_ERR_SOLVER, "Jacobi eigensolver did not converge in %d sweeps", max_sweeps);
  const double* Gfinal = wa.G[(sweep * (nb - 1)) & 1][0];  // the buffer the last round wrote
  return jacobi_finish(ctx, Gfinal, Vr, Vi, mp, m, lam_raw, perm, lam_out, Q_out);
}

// 4-wide-block driver (csrc/jacobi.cuh): comparator and fallback of the wide solver.
static int heev_jacobi(nls_ctx* ctx, const double* A, int m, double scale, double* lam_out, double* Q_out) {
  if (!ctx || !A || !lam_out || !Q_out || m < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_heev");
  CUDA_TRY(cudaSetDevice(ctx->device));
  int nb = (m + JB - 1) / JB;
  if (nb & 1) ++nb;
  if (nb < 2) nb = 2;
  const int mp = nb * JB, np = nb / 2;
  const size_t mm = (size_t)mp * mp;
  NLS_TRY(ensure(ctx, ctx->jac_mat, 4 * mm * 8));
  const int max_sweeps = 60;
  NLS_TRY(ensure(ctx, ctx->jac_small, (size_t)2 * np * 128 * 8 + (size_t)(2 * mp + 8) * 8 +
                                          (size_t)(2 * np + max_sweeps + 16 + mp) * 4));
  double* Gr = (double*)ctx->jac_mat.p;
  double* Gi = Gr + mm;
  double* Vr = Gi + mm;
  double* Vi = Vr + mm;
  double* Jbuf = (double*)ctx->jac_small.p;  // [2][np][128] (double-buffered by the persistent kernel)
  double* lam_raw = Jbuf + (size_t)2 * np * 128;
  double* fro2 = lam_raw + mp;
  int* flags = (int*)(fro2 + 8);            // [2][np]
  int* active = flags + 2 * np;             // [max_sweeps + 2]
  int* misc = active + max_sweeps + 2;      // [0] grid barrier counter, [1] sweeps executed
  int* perm = misc + 8;
  ProfScope scope(ctx, NLS_PROF_OTHER);
  jacobi_init_kernel<<<grid_for((long long)mm), 256, 0, ctx->stream>>>(A, m, mp, scale, Gr, Gi, Vr, Vi);
  NLS_TRY(check_launch(ctx, "jacobi_init_kernel"));
  sumsq_kernel<<<1, 1024, 0, ctx->stream>>>(Gr, Gi, (long long)mm, fro2);
  NLS_TRY(check_launch(ctx, "sumsq_kernel"));
  double h_fro2 = 0.0;
  CUDA_TRY(cudaMemcpyAsync(&h_fro2, fro2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  const double eps = 2.220446049250313e-16;
  double h_thr[2] = {eps * eps * h_fro2, 0.0};  // squared thresholds: absolute (eps ||G||_F), relative (off)
  double* thr = fro2 + 2;
  CUDA_TRY(cudaMemcpyAsync(thr, h_thr, 16, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  const long long tasks = (long long)np * (np + 1) / 2 + (long long)(mp / 8) * np;
  int sweep = 0, h_active = 1;
  const char* mode_env = getenv("NLS_JACOBI_MODE");  // "fused" or "split" (two kernels per round)
  const char* occ_env = getenv("NLS_JACOBI_OCC");    // resident CTAs per SM of the fused round kernel: 4, 5 or 6
  const int fused_occ = occ_env && atoi(occ_env) >= 4 && atoi(occ_env) <= 6 ? atoi(occ_env) : 4;
  const int fused_grid = ctx->sm_count * fused_occ;
  // Measured on B200: fused 58 ms vs split 70 ms at m = 1025, but 21 vs 19.5 ms at m = 513 (fewer tiles per
  // round to hide the pivot behind), so small problems keep the two-kernel round unless forced.
  const bool want_fused = mode_env ? strcmp(mode_env, "fused") == 0 : np >= 96;
  const bool fused = want_fused && np <= fused_grid && np <= 160;
  if (!ctx->jac_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->jac_stream, cudaStreamNonBlocking));
  if (fused) {
    // One kernel per round (priority tiles -> next pivots overlapped with the bulk update); one sweep
    // (nb - 1 identical launches; the round number lives on the device) is replayed as a CUDA graph.
    int* state = misc + 2;  // [0] round counter, [1..4] priority-tile counters, [5] CTA arrival counter
    CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)(2 * np + max_sweeps + 16) * 4, ctx->stream));
    JacobiArgs ja;
    ja.Gr = Gr; ja.Gi = Gi; ja.Vr = Vr; ja.Vi = Vi;
    ja.ld = mp; ja.nb = nb; ja.max_inner = ctx->jac_inner; ja.max_sweeps = max_sweeps;
    ja.thr = thr; ja.Jbuf = Jbuf; ja.flags = flags; ja.active = active;
    // J(0): pivots of the very first round.
    jacobi_pivot_kernel<<<(np + 3) / 4, 128, 0, ctx->stream>>>(Gr, Gi, mp, nb, 0, thr, ctx->jac_inner, Jbuf, flags, active);
    NLS_TRY(check_launch(ctx, "jacobi_pivot_kernel"));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (!ctx->jac_graph || ctx->jac_graph_key != (const void*)Gr || ctx->jac_graph_key2 != (const void*)ctx->jac_small.p || ctx->jac_graph_nb != -(nb * 8 + fused_occ)) {
      if (ctx->jac_graph) {
        cudaGraphExecDestroy(ctx->jac_graph);
        ctx->jac_graph = nullptr;
      }
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(ctx->jac_stream, cudaStreamCaptureModeRelaxed));
      for (int round = 0; round < nb - 1; ++round) {
        if (fused_occ == 4) jacobi_round_kernel<4><<<fused_grid, 256, 0, ctx->jac_stream>>>(ja, state);
        else if (fused_occ == 5) jacobi_round_kernel<5><<<fused_grid, 256, 0, ctx->jac_stream>>>(ja, state);
        else jacobi_round_kernel<6><<<fused_grid, 256, 0, ctx->jac_stream>>>(ja, state);
      }
      CUDA_TRY(cudaStreamEndCapture(ctx->jac_stream, &graph));
      cudaError_t ge = cudaGraphInstantiate(&ctx->jac_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ge != cudaSuccess) return fail(NLS_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ge));
      ctx->jac_graph_key = (const void*)Gr;
      ctx->jac_graph_key2 = (const void*)ctx->jac_small.p;
      ctx->jac_graph_nb = -(nb * 8 + fused_occ);  // negative: fused-round graph
    }
    for (; sweep < max_sweeps && h_active > 0; ++sweep) {
      CUDA_TRY(cudaGraphLaunch(ctx->jac_graph, ctx->jac_stream));
      ctx->launches += nb - 1;
      CUDA_TRY(cudaMemcpyAsync(&h_active, active + sweep, sizeof(int), cudaMemcpyDeviceToHost, ctx->jac_stream));
      CUDA_TRY(cudaStreamSynchronize(ctx->jac_stream));
    }
  } else {
    const int upd_grid = (int)std::min<long long>((tasks + 7) / 8, (long long)ctx->sm_count * 8);
    if (!ctx->jac_graph || ctx->jac_graph_key != (const void*)Gr || ctx->jac_graph_key2 != (const void*)ctx->jac_small.p || ctx->jac_graph_nb != nb) {
      // Capture one sweep: reset the rotation counter, then nb-1 rounds of (pivot, update).
      if (ctx->jac_graph) {
        cudaGraphExecDestroy(ctx->jac_graph);
        ctx->jac_graph = nullptr;
      }
      cudaGraph_t graph = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(ctx->jac_stream, cudaStreamCaptureModeRelaxed));
      cudaMemsetAsync(active, 0, sizeof(int), ctx->jac_stream);
      for (int round = 0; round < nb - 1; ++round) {
        jacobi_pivot_kernel<<<(np + 3) / 4, 128, 0, ctx->jac_stream>>>(Gr, Gi, mp, nb, round, thr, ctx->jac_inner, Jbuf,
                                                                      flags, active);
        jacobi_update_kernel<<<upd_grid, 256, 0, ctx->jac_stream>>>(Gr, Gi, Vr, Vi, mp, nb, round, Jbuf, flags);
      }
      CUDA_TRY(cudaStreamEndCapture(ctx->jac_stream, &graph));
      cudaError_t ge = cudaGraphInstantiate(&ctx->jac_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ge != cudaSuccess) return fail(NLS_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ge));
      ctx->jac_graph_key = (const void*)Gr;
      ctx->jac_graph_key2 = (const void*)ctx->jac_small.p;
      ctx->jac_graph_nb = nb;
    }
    for (; sweep < max_sweeps && h_active > 0; ++sweep) {
      CUDA_TRY(cudaGraphLaunch(ctx->jac_graph, ctx->jac_stream));
      ctx->launches += 2 * (nb - 1);
      CUDA_TRY(cudaMemcpyAsync(&h_active, active, sizeof(int), cudaMemcpyDeviceToHost, ctx->jac_stream));
      CUDA_TRY(cudaStreamSynchronize(ctx->jac_stream));
    }
  }
  ctx->eig_sweeps = sweep;
  if (h_active > 0) return fail(NLS_ERR_SOLVER, "Jacobi eigensolver did not converge in %d sweeps", max_sweeps);
  return jacobi_finish(ctx, Gr, Vr, Vi, mp, m, lam_raw, perm, lam_out, Q_out);
}

#include "eig_driver.inc"

// Hand-written tridiagonalisation + divide and conquer + back-transformation (csrc/hetrd.cuh, csrc/stedc.cuh).
static int heev_dc(nls_ctx* ctx, const double* A, int m, double scale, double* lam_out, double* Q_out) {
  if (!A || !lam_out || !Q_out || m < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_heev");
  double* Y = nullptr;
  long long ldy = 0, yplane = 0;
  NLS_TRY(heev_tridiag<true>(ctx, A, m, m, scale, lam_out, &Y, &ldy, &yplane));
  bt_export_complex_kernel<<<dim3((m + 31) / 32, (m + 31) / 32), 256, 0, ctx->stream>>>(Y, ldy, yplane, m, Q_out);
  return check_launch(ctx, "bt_export_complex_kernel");
}

// Eigen-decomposition of a real symmetric tridiagonal matrix given on the HOST (diagonal d, off-diagonal e):
// lam_out (device, n) ascending, Zt_out (device, n x n row-major): row k = eigenvector of lam[k].
extern "C" int nls_stedc(nls_ctx* ctx, int n, const double* d_host, const double* e_host, double* lam_out,
                         double* Zt_out) {
  if (!ctx || !d_host || !lam_out || !Zt_out || n < 1 || (n > 1 && !e_host)) return fail(NLS_ERR_INVALID, "bad argument to nls_stedc");
  CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<double> d(d_host, d_host + n), e(e_host, e_host + (n > 1 ? n - 1 : 0));
  return stedc_device(ctx, n, d, e, lam_out, Zt_out, n);
}

// The tridiagonal form (d: n, e: n - 1, host arrays) that the last nls_heev / dual eigensolve reduced its matrix to.
extern "C" int nls_ctx_last_tridiagonal(nls_ctx* ctx, int n, double* d_host, double* e_host) {
  if (!ctx || !d_host || n < 1 || (int)ctx->last_d.size() != n) return fail(NLS_ERR_INVALID, "no tridiagonal form of size %d", n);
  std::copy(ctx->last_d.begin(), ctx->last_d.end(), d_host);
  if (n > 1 && e_host) std::copy(ctx->last_e.begin(), ctx->last_e.end(), e_host);
  return NLS_OK;
}

extern "C" int nls_heev(nls_ctx* ctx, const double* A, int m, double scale, double* lam_out, double* Q_out) {
  if (!ctx) return fail(NLS_ERR_INVALID, "ctx is null");
  // auto: the tridiagonalisation + divide-and-conquer solver; the block-Jacobi kernels (kind 0) and cuSOLVER's Zheevd
  // (kind 1) stay selectable as comparators.
  ctx->eig_sweeps = 0;
  if (ctx->eig_kind == 2 || ctx->eig_kind == 3) return heev_dc(ctx, A, m, scale, lam_out, Q_out);
  if (ctx->eig_kind == 1) return heev_cusolver(ctx, A, m, scale, lam_out, Q_out);
  if (!A || !lam_out || !Q_out || m < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_heev");
  // Block width: 8 (16 x 16 pivots, default) or the original 4 (8 x 8 pivots; NLS_JACOBI_JB=4, all its variants).
  const char* jb_env = getenv("NLS_JACOBI_JB");
  const int jb = jb_env ? atoi(jb_env) : 8;
  if (jb == 8 && m <= 2552) return heev_jacobi_wide<8>(ctx, A, m, scale, lam_out, Q_out);
  return heev_jacobi(ctx, A, m, scale, lam_out, Q_out);
}

extern "C" int nls_ctx_set_eigensolver(nls_ctx* ctx, int kind) {
  if (!ctx || kind < 0 || kind > 3)
    return fail(NLS_ERR_INVALID, "eigensolver kind must be 0 (block Jacobi), 1 (cuSOLVER), 2 (auto) or 3 (divide and conquer)");
  ctx->eig_kind = kind;
  return NLS_OK;
}

extern "C" int nls_ctx_last_eig_sweeps(const nls_ctx* ctx) { return ctx ? ctx->eig_sweeps : 0; }

extern "C" int nls_ctx_set_gemm_core(nls_ctx* ctx, int kind) {
  if (!ctx || kind < 0 || kind > 2)
    return fail(NLS_ERR_INVALID, "GEMM core must be 0 (FP64 DMMA everywhere), 1 (Gram / projection / predict_std on the INT8 "
                                 "tensor cores, Ozaki scheme) or 2 (the γ sweep as well)");
  ctx->gemm_core = kind;
  return NLS_OK;
}

// v = Q^H b inv_c  and (optionally, gamma >= 0) beta_eig = Q (v / (lam + gamma)).
extern "C" int nls_primal_coeffs(nls_ctx* ctx, const double* Q, const double* lam, const double* b, int m,
                                 double inv_c, double gamma, double* v_out, double* beta_eig_out) {
  if (!ctx || !Q || !lam || !v_out || m < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_primal_coeffs");
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (b) {
    project_rhs_kernel<<<m, 256, 0, ctx->stream>>>(Q, b, m, inv_c, v_out);
    NLS_TRY(check_launch(ctx, "project_rhs_kernel"));
  }
  if (beta_eig_out) {
    eigen_beta_kernel<<<m, 256, 0, ctx->stream>>>(Q, v_out, lam, m, gamma, beta_eig_out);
    NLS_TRY(check_launch(ctx, "eigen_beta_kernel"));
  }
  return NLS_OK;
}

// Hand-written blocked Cholesky M = U^H U in place (csrc/potrf.cuh): M is n x n row-major with pitch ld elements
// (complex: interleaved); the upper triangle becomes U, the strict lower triangle is left as it was.
template <bool C>
static int chol_factor(nls_ctx* ctx, double* Mraw, int n, long long ld) {
  using T = typename HS<C>::T;
  T* M = reinterpret_cast<T*>(Mraw);
  const int planes = C ? 2 : 1;
  const long long npad = round_up(n, 128);
  NLS_TRY(ensure(ctx, ctx->eig.chol, (size_t)planes * npad * CNB * 8 + 64));
  double* PT = (double*)ctx->eig.chol.p;
  int* info = (int*)(PT + (size_t)planes * npad * CNB);
  cudaStream_t st = ctx->stream;
  CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(int), st));
  const size_t smem_d = sizeof(T) * CNB * (CNB + 1), smem_p = smem_d + sizeof(T) * CNB * CNB;
  static bool attr_set[64][2] = {{false}};
  if (!attr_set[ctx->device & 63][C]) {
    CUDA_TRY(cudaFuncSetAttribute(chol_diag_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
    CUDA_TRY(cudaFuncSetAttribute(chol_panel_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    attr_set[ctx->device & 63][C] = true;
  }
  ProfScope scope(ctx, NLS_PROF_OTHER);
  for (int k0 = 0; k0 < n; k0 += CNB) {
    const int nb = std::min(CNB, n - k0);
    chol_diag_kernel<C><<<1, 256, smem_d, st>>>(M, ld, k0, nb, info);
    NLS_TRY(check_launch(ctx, "chol_diag_kernel"));
    const int r0 = k0 + nb, nt = n - r0;
    if (nt <= 0) break;
    chol_panel_kernel<C><<<(nt + 63) / 64, 64, smem_p, st>>>(M, ld, n, k0, nb, PT, npad * CNB);
    NLS_TRY(check_launch(ctx, "chol_panel_kernel"));
    typename OpCholUpdate<C>::Params up;
    up.A = Operand{PT + (long long)r0 * CNB, CNB, nt, CNB, (int)npad, 0};
    up.B = up.A;
    up.nt = nt;
    up.Mout = reinterpret_cast<double*>(M + (long long)r0 * ld + r0);
    up.ld = ld;
    const dim3 grid((nt + BN - 1) / BN, (nt + BM - 1) / BM);
    if (C) {
      NLS_TRY((launch_gemm<MODE_COMPLEX, OpCholUpdate<C>>(ctx, up, grid, planes * npad - r0, planes * npad - r0, NLS_PROF_OTHER, "chol_update")));
    } else {
      NLS_TRY((launch_gemm<MODE_REAL, OpCholUpdate<C>>(ctx, up, grid, planes * npad - r0, planes * npad - r0, NLS_PROF_OTHER, "chol_update")));
    }
  }
  int h_info = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h_info != 0) return fail(NLS_ERR_SOLVER, "Cholesky factorisation failed: the leading minor of order %d is not positive definite", h_info);
  return NLS_OK;
}

// v <- (U^H U)^-1 v by block forward and back substitution.
template <bool C>
static int chol_solve(nls_ctx* ctx, const double* Uraw, int n, long long ld, double* vraw) {
  using T = typename HS<C>::T;
  const T* U = reinterpret_cast<const T*>(Uraw);
  T* v = reinterpret_cast<T*>(vraw);
  cudaStream_t st = ctx->stream;
  for (int k0 = 0; k0 < n; k0 += CNB) {
    const int nb = std::min(CNB, n - k0), nt = n - k0 - nb;
    chol_solve_diag_kernel<C><<<1, 64, 0, st>>>(U, ld, k0, nb, 0, v);
    if (nt > 0) chol_solve_update_kernel<C><<<(nt + 255) / 256, 256, 0, st>>>(U, ld, n, k0, nb, 0, v);
    ctx->launches += 1 + (nt > 0);
  }
  for (int k0 = ((n - 1) / CNB) * CNB; k0 >= 0; k0 -= CNB) {
    const int nb = std::min(CNB, n - k0);
    chol_solve_diag_kernel<C><<<1, 64, 0, st>>>(U, ld, k0, nb, 1, v);
    if (k0 > 0) chol_solve_update_kernel<C><<<(k0 + 255) / 256, 256, 0, st>>>(U, ld, n, k0, nb, 1, v);
    ctx->launches += 1 + (k0 > 0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(NLS_ERR_CUDA, "triangular solve launch failed: %s", cudaGetErrorString(e));
  return NLS_OK;
}

// B_out = U^-1 (n x n row-major, pitch n) of the upper triangle of U (pitch ld elements); csrc/potrf.cuh.
template <bool C>
static int tri_inverse(nls_ctx* ctx, const double* Uraw, int n, long long ld, double* Bout) {
  using T = typename HS<C>::T;
  const int planes = C ? 2 : 1;
  const long long npad = round_up(n, 128), ldu = round_up(n, 16);
  const size_t plane = (size_t)npad * ldu;
  EigBuffers& B = ctx->eig;
  NLS_TRY(ensure(ctx, B.qa, (size_t)planes * plane * 8));                  // planar U
  NLS_TRY(ensure(ctx, B.qb, (size_t)planes * plane * 8));                  // Xt
  NLS_TRY(ensure(ctx, B.qp, (size_t)planes * CNB * ldu * 8));              // R
  double* Up = (double*)B.qa.p;
  double* Xt = (double*)B.qb.p;
  double* R = (double*)B.qp.p;
  cudaStream_t st = ctx->stream;
  ProfScope scope(ctx, NLS_PROF_OTHER);
  CUDA_TRY(cudaMemsetAsync(Up, 0, (size_t)planes * plane * 8, st));
  CUDA_TRY(cudaMemsetAsync(Xt, 0, (size_t)planes * plane * 8, st));
  trtri_split_kernel<C><<<grid_for((long long)n * n), 256, 0, st>>>(reinterpret_cast<const T*>(Uraw), ld, n, Up, ldu, (long long)plane);
  NLS_TRY(check_launch(ctx, "trtri_split_kernel"));
  const size_t smem = sizeof(T) * CNB * (CNB + 1) + sizeof(T) * CNB * CNB;
  static bool attr_set[64][2] = {{false}};
  if (!attr_set[ctx->device & 63][C]) {
    CUDA_TRY(cudaFuncSetAttribute(trtri_block_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[ctx->device & 63][C] = true;
  }
  for (int k0 = ((n - 1) / CNB) * CNB; k0 >= 0; k0 -= CNB) {
    const int nb = std::min(CNB, n - k0), k_lo = k0 + nb;
    if (k_lo < n) {
      typename OpTrtriProduct<C>::Params pp;
      pp.A = Operand{Up + (long long)k0 * ldu, ldu, nb, n, (int)npad, 0};
      pp.B = Operand{Xt, ldu, n, n, (int)npad, 0};
      pp.nb = nb;
      pp.j0 = k_lo;
      pp.n = n;
      pp.k_lo = k_lo;
      pp.R = R;
      pp.ldr = ldu;
      pp.rplane = (long long)CNB * ldu;
      const dim3 grid((n - k_lo + BN - 1) / BN, 1);
      if (C) {
        NLS_TRY((launch_gemm<MODE_COMPLEX, OpTrtriProduct<C>>(ctx, pp, grid, planes * npad - k0, planes * npad, NLS_PROF_OTHER, "trtri_product")));
      } else {
        NLS_TRY((launch_gemm<MODE_REAL, OpTrtriProduct<C>>(ctx, pp, grid, planes * npad - k0, planes * npad, NLS_PROF_OTHER, "trtri_product")));
      }
    }
    trtri_block_kernel<C><<<(n - k0 + 63) / 64, 64, smem, st>>>(Up, ldu, (long long)plane, n, k0, nb, R, ldu, (long long)CNB * ldu, Xt);
    NLS_TRY(check_launch(ctx, "trtri_block_kernel"));
  }
  trtri_export_kernel<C><<<dim3((n + 31) / 32, (n + 31) / 32), 256, 0, st>>>(Xt, ldu, (long long)plane, n, reinterpret_cast<T*>(Bout));
  return check_launch(ctx, "trtri_export_kernel");
}

// U^-1 of an upper-triangular factor (anything below the diagonal of U is ignored): the basis of predict_std,
// (gamma C + A)^-1 = U^-1 U^-H (_neo_ls_svm.py:467-469 primal / :473-475 dual).  is_complex selects complex128.
extern "C" int nls_triangular_inverse(nls_ctx* ctx, const double* U, int n, int is_complex, double* B_out) {
  if (!ctx || !U || !B_out || n < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_triangular_inverse");
  CUDA_TRY(cudaSetDevice(ctx->device));
  return is_complex ? tri_inverse<true>(ctx, U, n, n, B_out) : tri_inverse<false>(ctx, U, n, n, B_out);
}

// U (upper, row-major, M = U^H U like scipy.linalg.cho_factor) of M = A + diag_shift * I, and
// beta = M^-1 b.   _neo_ls_svm.py:177-178.  Hand-written blocked Cholesky (csrc/potrf.cuh); cuSOLVER's Zpotrf / Zpotrs
// only as the comparator (nls_ctx_set_eigensolver(1) / NLS_EIG=cusolver).
extern "C" int nls_cholesky_solve(nls_ctx* ctx, const double* A, int m, double diag_shift, const double* b,
                                  double* U_out, double* beta_out) {
  if (!ctx || !A || !U_out || m < 1) return fail(NLS_ERR_INVALID, "bad argument to nls_cholesky_solve");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const long long mm = (long long)m * m;
  CUDA_TRY(cudaMemcpyAsync(U_out, A, (size_t)mm * 16, cudaMemcpyDeviceToDevice, ctx->stream));
  add_diag_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(U_out, m, diag_shift);
  NLS_TRY(check_launch(ctx, "add_diag_kernel"));
  if (ctx->eig_kind != 1) {
    NLS_TRY(chol_factor<true>(ctx, U_out, m, m));
    if (b && beta_out) {
      CUDA_TRY(cudaMemcpyAsync(beta_out, b, (size_t)m * 16, cudaMemcpyDeviceToDevice, ctx->stream));
      NLS_TRY(chol_solve<true>(ctx, U_out, m, m, beta_out));
    }
    return NLS_OK;
  }
  // Row-major M read column-major is conj(M); its LOWER Cholesky factor stored column-major is
  // exactly U row-major with M = U^H U.
  int lwork = 0;
  SOLVER_TRY(cusolverDnZpotrf_bufferSize(ctx->solver, CUBLAS_FILL_MODE_LOWER, m, (cuDoubleComplex*)U_out, m, &lwork));
  NLS_TRY(ensure(ctx, ctx->solver_ws, (size_t)lwork * 16 + 64));
  int* info = (int*)((char*)ctx->solver_ws.p + (size_t)lwork * 16);
  SOLVER_TRY(cusolverDnZpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, m, (cuDoubleComplex*)U_out, m,
                              (cuDoubleComplex*)ctx->solver_ws.p, lwork, info));
  int h_info = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (h_info != 0) return fail(NLS_ERR_SOLVER, "Cholesky factorisation failed: info = %d", h_info);
  if (b && beta_out) {
    // conj(M) conj(beta) = conj(b)
    CUDA_TRY(cudaMemcpyAsync(beta_out, b, (size_t)m * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    conj_vec_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(beta_out, m);
    NLS_TRY(check_launch(ctx, "conj_vec_kernel"));
    SOLVER_TRY(cusolverDnZpotrs(ctx->solver, CUBLAS_FILL_MODE_LOWER, m, 1, (cuDoubleComplex*)U_out, m,
                                (cuDoubleComplex*)beta_out, m, info));
    conj_vec_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(beta_out, m);
    NLS_TRY(check_launch(ctx, "conj_vec_kernel"));
  }
  return NLS_OK;
}

// ---------------------------------------------------------------------------------------------
// Projection on the INT8 tensor cores (csrc/ozaki.cuh)
// ---------------------------------------------------------------------------------------------
struct OzSweep {
  int G, m, ngrp, rb_per_group, row_blocks, nks;
  long long ldk;
  const double *d, *inv_d, *ascale;
  const int8_t* a_planes;
  int* ex_pu;
  double *pscale, *uscale;
  double *qscale, *gscale;  // [ngrp][2] global quantisation / recombination scales of P', U' (fused projection -> planes)
};

struct OzBasis {
  const int8_t* planes;    // [n_tiles][nks][S][64 x 32 B]
  const double* colscale;  // 2^(eA + eB_j) per complex column
  int n_tiles, nks, eA;
};

// The projection-type products run on the INT8 core when the context says so and K fits the INT32 accumulators.
static bool oz_usable(const nls_ctx* ctx, const MapGeom& g) {
  return ctx->gemm_core >= 1 && oz::feature_ksteps(g.D) * oz::KS <= oz::MAX_K && tail_split(g.m) > 0;
}

static int oz_attr(nls_ctx* ctx) {
  if (ctx->oz_attr) return NLS_OK;
  CUDA_TRY(cudaFuncSetAttribute(oz::gemm_kernel_i8<oz::IMAGE, oz::EpiVariance>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                oz::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(oz::gemm_kernel_i8<oz::IMAGE, oz::EpiSweep>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                oz::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(oz::gemm_kernel_i8<oz::IMAGE, oz::EpiProjectPlanes>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                oz::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(oz::gemm_kernel_i8<oz::IMAGE, oz::EpiProject>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                oz::SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(oz::gemm_kernel_i8<oz::IMAGE, oz::EpiGram>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                oz::SMEM_BYTES));
  ctx->oz_attr = true;
  return NLS_OK;
}

// Digit planes of the basis [Re Q^T ; Im Q^T] (bs.bt), once per sweep call.
static int oz_prep_basis(nls_ctx* ctx, const MapGeom& g, const BasisScratch& bs, int cols, OzBasis* out) {
  NLS_TRY(oz_attr(ctx));
  const int n_tiles = (cols + oz::TN / 2 - 1) / (oz::TN / 2);
  const int nks = oz::feature_ksteps(g.D);
  const int padded = n_tiles * (oz::TN / 2);
  NLS_TRY(ensure(ctx, ctx->oz_b, (size_t)n_tiles * nks * oz::B_STAGE));
  NLS_TRY(ensure(ctx, ctx->oz_small, (size_t)padded * 16));
  double* colscale = (double*)ctx->oz_small.p;
  int* ex = (int*)(colscale + padded);
  const int eA = oz::scale_exponent(1.0 / sqrt((double)g.D));  // |cos| D^-1/2, |sin| D^-1/2 <= D^-1/2
  ProfScope scope(ctx, NLS_PROF_SLICE);
  oz::basis_exponent_kernel<<<(padded + 7) / 8, 256, 0, ctx->stream>>>(bs.bt, g.Np, g.Dp, g.D, cols, padded, eA, ex, colscale);
  NLS_TRY(check_launch(ctx, "oz::basis_exponent_kernel"));
  oz::slice_basis_kernel<oz::IMAGE><<<grid_for((long long)n_tiles * nks * oz::TN * 2), 256, 0, ctx->stream>>>(
      bs.bt, g.Np, g.Dp, g.D, cols, ex, nks, n_tiles, (int8_t*)ctx->oz_b.p);
  NLS_TRY(check_launch(ctx, "oz::slice_basis_kernel"));
  out->planes = (const int8_t*)ctx->oz_b.p;
  out->colscale = colscale;
  out->n_tiles = n_tiles;
  out->nks = nks;
  out->eA = eA;
  return NLS_OK;
}

// Digit planes of the planar feature chunk in ctx->psi (the A operand of the projection-type products).
static int oz_slice_chunk(nls_ctx* ctx, const MapGeom& g, const OzBasis& ob, int rows) {
  const int row_blocks = (rows + oz::TM - 1) / oz::TM;
  NLS_TRY(ensure(ctx, ctx->oz_a, (size_t)((ctx->chunk_rows + oz::TM - 1) / oz::TM) * ob.nks * oz::A_STAGE));
  ProfScope slice_scope(ctx, NLS_PROF_SLICE);
  oz::slice_rows_kernel<oz::IMAGE><<<grid_for((long long)row_blocks * ob.nks * oz::TM * 2), 256, 0, ctx->stream>>>(
      (const double*)ctx->psi.p, 2LL * g.Dp, rows, g.D, g.Dp, ldexp(1.0, oz::FRAC_BITS - ob.eA), ob.nks, row_blocks,
      (int8_t*)ctx->oz_a.p);
  return check_launch(ctx, "oz::slice_rows_kernel");
}

// P = Re(T v), U = |T|^2 / c for the first `cols` columns of T = phi Q, phi = the planar chunk in ctx->psi.
static int oz_project_chunk(nls_ctx* ctx, const MapGeom& g, const BasisScratch& bs, const OzBasis& ob, int rows, int cols,
                            double inv_c, double* P, double* U, const OzSweep* sw = nullptr) {
  const int row_blocks = (rows + oz::TM - 1) / oz::TM;
  NLS_TRY(oz_slice_chunk(ctx, g, ob, rows));
  ProfScope scope(ctx, NLS_PROF_PROJECT);
  if (sw) {  // straight into the digit planes of the INT8 sweep's B operand (no FP64 P, U)
    oz::GemmParams gp{(const int8_t*)ctx->oz_a.p, ob.planes, ob.nks, row_blocks, ob.n_tiles, 0, row_blocks * ob.n_tiles, 1, ob.nks, 0, 1, 1 << 30, 0};
    oz::EpiProjectPlanes::Params ep{rows, cols, ob.colscale, bs.bias_r, bs.bias_i, bs.v_r, bs.v_i, inv_c, sw->inv_d, sw->qscale,
                                    sw->ngrp, sw->nks, sw->ldk, (int8_t*)ctx->oz_sb.p};
    const int grid = (int)std::min<long long>((long long)row_blocks * ob.n_tiles, ctx->sm_count);
    oz::gemm_kernel_i8<oz::IMAGE, oz::EpiProjectPlanes><<<grid, oz::THREADS, oz::SMEM_BYTES, ctx->stream>>>(gp, ep);
    return check_launch(ctx, "oz::gemm_kernel_i8<EpiProjectPlanes>");
  }
  oz::GemmParams gp{(const int8_t*)ctx->oz_a.p, ob.planes, ob.nks, row_blocks, ob.n_tiles, 0, row_blocks * ob.n_tiles, 1, ob.nks, 0, 1, 1 << 30, 0};
  oz::EpiProject::Params ep{rows, cols, ob.colscale, bs.bias_r, bs.bias_i, bs.v_r, bs.v_i, inv_c, P, U, g.ldp};
  const int grid = (int)std::min<long long>((long long)row_blocks * ob.n_tiles, ctx->sm_count);
  oz::gemm_kernel_i8<oz::IMAGE, oz::EpiProject><<<grid, oz::THREADS, oz::SMEM_BYTES, ctx->stream>>>(gp, ep);
  return check_launch(ctx, "oz::gemm_kernel_i8");
}

// γ sweep on the INT8 core: operands that do not depend on the chunk (column scalings, digit planes of r').
static_assert(oz::MAX_SWEEP_GROUPS >= 2, "");
constexpr int OZ_SWEEP_GROUPS = 2;  // γ groups with their own column scaling (scripts/sweep_int8_study.py: 1 already suffices)

static int oz_prep_sweep(nls_ctx* ctx, const MapGeom& g, const double* gammas, const double* lam, int G, const double* v_r,
                         const double* v_i, double inv_c, OzSweep* out) {
  NLS_TRY(oz_attr(ctx));
  OzSweep o{};
  o.G = G;
  o.m = g.m;
  o.row_blocks = (G + oz::TM - 1) / oz::TM;
  o.rb_per_group = (o.row_blocks + OZ_SWEEP_GROUPS - 1) / OZ_SWEEP_GROUPS;
  o.ngrp = (o.row_blocks + o.rb_per_group - 1) / o.rb_per_group;
  o.nks = (g.m + oz::KS - 1) / oz::KS;
  o.ldk = (long long)o.nks * oz::KS;
  const long long cap = ctx->chunk_rows;
  // small arrays: d, inv_d [ngrp][ldk]; ascale [G]; pscale, uscale [cap][ngrp]; exA [G]; ex_pu [cap][ngrp][2]
  const size_t n_d = (size_t)o.ngrp * o.ldk, n_g = (size_t)round_up(G, 2), n_pu = (size_t)cap * o.ngrp;
  NLS_TRY(ensure(ctx, ctx->oz_ssmall, (2 * n_d + n_g + 2 * n_pu + 4 * (size_t)o.ngrp) * 8 + (n_g + 2 * n_pu) * 4));
  double* base = (double*)ctx->oz_ssmall.p;
  o.qscale = base;
  o.gscale = base + 2 * o.ngrp;
  double* d = base + 4 * o.ngrp;
  double* inv_d = d + n_d;
  double* ascale = inv_d + n_d;
  o.pscale = ascale + n_g;
  o.uscale = o.pscale + n_pu;
  int* exA = (int*)(o.uscale + n_pu);
  o.ex_pu = exA + n_g;
  NLS_TRY(ensure(ctx, ctx->oz_sa, (size_t)o.row_blocks * o.nks * oz::A_STAGE));
  NLS_TRY(ensure(ctx, ctx->oz_sb, (size_t)((cap + oz::TN / 2 - 1) / (oz::TN / 2)) * o.ngrp * o.nks * oz::B_STAGE));
  ProfScope scope(ctx, NLS_PROF_SLICE);
  oz::sweep_groups_kernel<<<o.ngrp, 256, 0, ctx->stream>>>(gammas, lam, G, g.m, o.ngrp, o.rb_per_group, o.ldk, d, inv_d);
  NLS_TRY(check_launch(ctx, "oz::sweep_groups_kernel"));
  oz::rgamma_exponent_kernel<<<(G + 7) / 8, 256, 0, ctx->stream>>>(gammas, lam, d, G, g.m, o.rb_per_group, o.ldk, exA, ascale);
  NLS_TRY(check_launch(ctx, "oz::rgamma_exponent_kernel"));
  oz::slice_rgamma_kernel<oz::IMAGE><<<grid_for((long long)o.row_blocks * o.nks * oz::TM * 2), 256, 0, ctx->stream>>>(
      gammas, lam, d, exA, G, g.m, o.rb_per_group, o.ldk, o.nks, o.row_blocks, (int8_t*)ctx->oz_sa.p);
  NLS_TRY(check_launch(ctx, "oz::slice_rgamma_kernel"));
  oz::sweep_bounds_kernel<<<o.ngrp, 256, 0, ctx->stream>>>(v_r, v_i, inv_d, g.m, o.ldk, inv_c, o.qscale, o.gscale);
  NLS_TRY(check_launch(ctx, "oz::sweep_bounds_kernel"));
  o.d = d;
  o.inv_d = inv_d;
  o.ascale = ascale;
  o.a_planes = (const int8_t*)ctx->oz_sa.p;
  *out = o;
  return NLS_OK;
}

// num / den / LOO error sums (and the σ² stash) of one chunk from its P, U (row pitch ldp).
// `planes_from`: >= 0 when the projection epilogue already wrote the planes of the k steps below it with the global scales
// (then only the spill columns' k steps are sliced here, from the FP64 P, U the spill kernels wrote); -1: everything from P, U.
static int oz_sweep_chunk(nls_ctx* ctx, const MapGeom& g, const OzSweep& o, int rows, const double* P, const double* U,
                          const double* y, const double* s, int is_classifier, double* den_out, double* sums_out,
                          int planes_from = -1) {
  const int n_tiles = (rows + oz::TN / 2 - 1) / (oz::TN / 2);
  NLS_TRY(ensure(ctx, ctx->part, (size_t)2 * ((ctx->chunk_rows + oz::TN / 2 - 1) / (oz::TN / 2)) * 3 * o.G * 8));
  if (planes_from >= 0) {
    if (planes_from < o.nks) {
      ProfScope scope(ctx, NLS_PROF_SLICE);
      oz::slice_pu_kernel<oz::IMAGE><<<grid_for((long long)n_tiles * o.ngrp * (o.nks - planes_from) * oz::TN * 2), 256, 0, ctx->stream>>>(
          P, U, g.ldp, rows, g.m, o.inv_d, nullptr, o.qscale, o.ngrp, o.ldk, planes_from, o.nks, n_tiles, (int8_t*)ctx->oz_sb.p);
      NLS_TRY(check_launch(ctx, "oz::slice_pu_kernel"));
    }
  } else {
    ProfScope scope(ctx, NLS_PROF_SLICE);
    oz::pu_exponent_kernel<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(P, U, g.ldp, rows, g.m, o.inv_d, o.ngrp, o.ldk, o.ex_pu, o.pscale,
                                                                  o.uscale);
    NLS_TRY(check_launch(ctx, "oz::pu_exponent_kernel"));
    oz::slice_pu_kernel<oz::IMAGE><<<grid_for((long long)n_tiles * o.ngrp * o.nks * oz::TN * 2), 256, 0, ctx->stream>>>(
        P, U, g.ldp, rows, g.m, o.inv_d, o.ex_pu, nullptr, o.ngrp, o.ldk, 0, o.nks, n_tiles, (int8_t*)ctx->oz_sb.p);
    NLS_TRY(check_launch(ctx, "oz::slice_pu_kernel"));
  }
  {
    ProfScope scope(ctx, NLS_PROF_SWEEP);
    oz::GemmParams gp{o.a_planes, (const int8_t*)ctx->oz_sb.p, o.nks, o.row_blocks, n_tiles, 0, o.row_blocks * n_tiles, 1, o.nks, 0,
                      o.ngrp, o.rb_per_group, 1};
    oz::EpiSweep::Params ep{rows, o.G, o.ngrp, o.rb_per_group, o.ascale, o.pscale, o.uscale, planes_from >= 0 ? o.gscale : nullptr,
                            y, s, is_classifier, (double*)ctx->part.p, den_out, (long long)o.G};
    const int grid = (int)std::min<long long>((long long)gp.tiles, ctx->sm_count);
    oz::gemm_kernel_i8<oz::IMAGE, oz::EpiSweep><<<grid, oz::THREADS, oz::SMEM_BYTES, ctx->stream>>>(gp, ep);
    NLS_TRY(check_launch(ctx, "oz::gemm_kernel_i8<EpiSweep>"));
  }
  sweep_reduce_wide_kernel<<<(3 * o.G + 31) / 32, dim3(32, 32), 0, ctx->stream>>>((const double*)ctx->part.p, 2 * n_tiles, o.G, sums_out);
  return check_launch(ctx, "sweep_reduce_wide_kernel");
}

// ---------------------------------------------------------------------------------------------
// Stage 4a + 4b
// ---------------------------------------------------------------------------------------------
extern "C" int nls_primal_loo_sweep(nls_ctx* ctx, const double* X, const double* y, const double* s, int64_t n, int d,
                                    const double* shift, const double* W, int D, const double* Q, const double* lam,
                                    const double* v, double inv_c, const double* gammas, int G, int is_classifier,
                                    double* sums_out, double* sigma2_stash) {
  NLS_TRY(check_map_args(ctx, X, n, d, shift, W, D));
  if (!y || !s || !Q || !lam || !v || !gammas || !sums_out || G < 1) return fail(NLS_ERR_INVALID, "null pointer");
  const MapGeom g = geom(d, D);
  NLS_TRY(prep_weights(ctx, g, W));
  BasisScratch bs;
  NLS_TRY(prep_basis(ctx, g, Q, &bs));
  split_complex_kernel<<<(g.m + 255) / 256, 256, 0, ctx->stream>>>(v, g.m, bs.v_r, bs.v_i);
  NLS_TRY(check_launch(ctx, "split_complex_kernel"));
  const long long cap = ctx->chunk_rows;
  NLS_TRY(ensure(ctx, ctx->psi, (size_t)cap * 2 * g.Dp * 8));
  NLS_TRY(ensure(ctx, ctx->pu, (size_t)2 * cap * g.ldp * 8));
  NLS_TRY(ensure(ctx, ctx->rt, (size_t)round_up(G, BN) * g.ldp * 8));
  const int max_mtiles = (int)((cap + BM - 1) / BM);
  NLS_TRY(ensure(ctx, ctx->part, (size_t)max_mtiles * 3 * G * 8));
  build_rt_kernel<<<grid_for((long long)G * g.ldp), 256, 0, ctx->stream>>>(gammas, lam, G, g.m, g.ldp,
                                                                          (double*)ctx->rt.p);
  NLS_TRY(check_launch(ctx, "build_rt_kernel"));
  CUDA_TRY(cudaMemsetAsync(sums_out, 0, (size_t)3 * G * 8, ctx->stream));
  double* P = (double*)ctx->pu.p;
  double* U = P + cap * g.ldp;
  // The projection T = phi Q runs on the INT8 tensor cores (Ozaki scheme, csrc/ozaki.cuh) unless the context says DMMA.
  const bool use_oz = oz_usable(ctx, g);
  OzBasis oz{};
  if (use_oz) NLS_TRY(oz_prep_basis(ctx, g, bs, tail_split(g.m), &oz));
  const bool use_oz_sweep = ctx->gemm_core >= 2 && g.m <= oz::MAX_K;
  OzSweep ozs{};
  if (use_oz_sweep) NLS_TRY(oz_prep_sweep(ctx, g, gammas, lam, G, bs.v_r, bs.v_i, inv_c, &ozs));
  // projection and sweep both on the INT8 core: the projection epilogue writes the sweep's operand planes itself
  const char* fuse_env = getenv("NLS_OZ_FUSE");  // NLS_OZ_FUSE=0: planes sliced from FP64 P, U instead (A/B timing)
  const bool fused_planes = use_oz && use_oz_sweep && tail_split(g.m) % (oz::TN / 2) == 0 && !(fuse_env && fuse_env[0] == '0');
  for (int64_t i0 = 0; i0 < n; i0 += cap) {
    const int rows = (int)std::min<int64_t>(cap, n - i0);
    const int mtiles = (rows + BM - 1) / BM;
    // One spill column (m = D + 1, D a multiple of 64): its Re T / Im T are row dot products fused into the feature-map
    // epilogue, so the chunk is not read a second time by a GEMV tail.
    const int full_cols = tail_split(g.m);  // columns handled by 64-wide tiles
    const bool fused_spill = g.m - full_cols == 1;
    RowDots dots{};
    if (fused_spill) {
      const double* qr = bs.bt + (size_t)full_cols * g.Dp;
      const double* qi = bs.bt + (size_t)(g.Np + full_cols) * g.Dp;
      dots.x[0] = qr; dots.y[0] = qi; dots.ysign[0] = 1.0;   // Re((c - i s)(a + i b)) = c a + s b
      dots.x[1] = qi; dots.y[1] = qr; dots.ysign[1] = -1.0;  // Im                     = c b - s a
      dots.stride = 1; dots.count = 2; dots.ld = cap;
      NLS_TRY(rowdots_buffer(ctx, g, &dots.out));
    }
    NLS_TRY(feature_chunk(ctx, g, X + i0 * d, shift, rows, FM_PLANAR, (double*)ctx->psi.p, 2LL * g.Dp, g.Dp, nullptr,
                          fused_spill ? &dots : nullptr));
    if (use_oz) {
      NLS_TRY(oz_project_chunk(ctx, g, bs, oz, rows, full_cols, inv_c, P, U, fused_planes ? &ozs : nullptr));
    } else {
    OpProject::Params pp;
    pp.A = psi_operand(ctx, g, rows);
    pp.B = basis_operand(g, bs.bt);
    pp.n_rows = rows;
    pp.m = g.m;
    pp.bias_r = bs.bias_r;
    pp.bias_i = bs.bias_i;
    pp.v_r = bs.v_r;
    pp.v_i = bs.v_i;
    pp.inv_c = inv_c;
    pp.P = P;
    pp.U = U;
    pp.ld = g.ldp;
    pp.m = full_cols;
    NLS_TRY((launch_gemm<MODE_COMPLEX, OpProject>(ctx, pp, dim3((full_cols + BN - 1) / BN, mtiles), rows, 2 * g.Np,
                                                   NLS_PROF_PROJECT, "project")));
    }
    if (fused_spill) {
      ProfScope scope(ctx, NLS_PROF_PROJECT);
      project_spill_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(dots.out, (g.D + BN - 1) / BN, cap, rows, full_cols,
                                                                       bs.bias_r, bs.bias_i, bs.v_r, bs.v_i, inv_c, P, U,
                                                                       g.ldp);
      NLS_TRY(check_launch(ctx, "project_spill_kernel"));
    } else if (full_cols < g.m) {
      ProfScope scope(ctx, NLS_PROF_PROJECT);
      project_tail_kernel<<<(rows * 32 + 255) / 256, 256, 0, ctx->stream>>>(
          (const double*)ctx->psi.p, 2LL * g.Dp, g.Dp, rows, D, bs.bt, g.Dp, g.Np, full_cols, g.m, bs.bias_r, bs.bias_i, 0,
          bs.v_r, bs.v_i, inv_c, P, U, g.ldp, nullptr, nullptr);
      NLS_TRY(check_launch(ctx, "project_tail_kernel"));
    }
    if (use_oz_sweep) {
      NLS_TRY(oz_sweep_chunk(ctx, g, ozs, rows, P, U, y + i0, s + i0, is_classifier, sigma2_stash ? sigma2_stash + i0 * G : nullptr,
                             sums_out, fused_planes ? full_cols / oz::KS : -1));
      continue;
    }
    OpSweep::Params sp;
    sp.A = Operand{P, g.ldp, (int)(cap + rows), g.m, (int)cap, 0};
    sp.B = Operand{(const double*)ctx->rt.p, g.ldp, G, g.m, 0, 0};
    sp.n_rows = rows;
    sp.G = G;
    sp.y = y + i0;
    sp.s = s + i0;
    sp.is_classifier = is_classifier;
    sp.part = (double*)ctx->part.p;
    sp.den_out = sigma2_stash ? sigma2_stash + i0 * G : nullptr;
    sp.den_ld = G;
    NLS_TRY((launch_gemm<MODE_DUAL_A, OpSweep>(ctx, sp, dim3((G + BN - 1) / BN, mtiles), cap + rows, G,
                                                NLS_PROF_SWEEP, "sweep")));
    sweep_reduce_kernel<<<(3 * G + 255) / 256, 256, 0, ctx->stream>>>((const double*)ctx->part.p, mtiles, G, sums_out);
    NLS_TRY(check_launch(ctx, "sweep_reduce_kernel"));
  }
  return NLS_OK;
}

// ---------------------------------------------------------------------------------------------
// Shared by stage 4c and stage 5: per-row yhat (two coefficient vectors) and sigma2 for a chunk.
// ---------------------------------------------------------------------------------------------
static int variance_chunk(nls_ctx* ctx, const MapGeom& g, const BasisScratch& bs, int rows, double* sigma2_out,
                          int b_upper = 0, const OzBasis* ob = nullptr) {
  const int full_cols = tail_split(g.m);
  const long long cap = ctx->chunk_rows;
  if (ob) {
    // INT8 core: two partial sums (column halves) per 32-column tile, the spill columns after them.
    const int ntiles = 2 * ob->n_tiles + (full_cols < g.m ? 1 : 0);
    NLS_TRY(ensure(ctx, ctx->part, (size_t)ntiles * cap * 8));
    NLS_TRY(oz_slice_chunk(ctx, g, *ob, rows));
    const int row_blocks = (rows + oz::TM - 1) / oz::TM;
    {
      ProfScope scope(ctx, NLS_PROF_VARIANCE);
      oz::GemmParams gp{(const int8_t*)ctx->oz_a.p, ob->planes, ob->nks, row_blocks, ob->n_tiles, 0, row_blocks * ob->n_tiles,
                        1, ob->nks, b_upper ? 1 : 0, 1, 1 << 30, 0};
      oz::EpiVariance::Params ep{rows, full_cols, ob->colscale, bs.bias_r, bs.bias_i, bs.w, (double*)ctx->part.p, cap};
      const int grid = (int)std::min<long long>((long long)row_blocks * ob->n_tiles, ctx->sm_count);
      oz::gemm_kernel_i8<oz::IMAGE, oz::EpiVariance><<<grid, oz::THREADS, oz::SMEM_BYTES, ctx->stream>>>(gp, ep);
      NLS_TRY(check_launch(ctx, "oz::gemm_kernel_i8<EpiVariance>"));
      if (full_cols < g.m) {
        project_tail_kernel<<<(rows * 32 + 255) / 256, 256, 0, ctx->stream>>>(
            (const double*)ctx->psi.p, 2LL * g.Dp, g.Dp, rows, g.D, bs.bt, g.Dp, g.Np, full_cols, g.m, bs.bias_r, bs.bias_i, 1,
            nullptr, nullptr, 0.0, nullptr, nullptr, 0, bs.w, (double*)ctx->part.p + (size_t)2 * ob->n_tiles * cap);
        NLS_TRY(check_launch(ctx, "project_tail_kernel"));
      }
    }
    rowsum_reduce_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>((const double*)ctx->part.p, ntiles, cap, rows, sigma2_out);
    return check_launch(ctx, "rowsum_reduce_kernel");
  }
  const int gemm_tiles = (full_cols + BN - 1) / BN;
  const int ntiles = gemm_tiles + (full_cols < g.m ? 1 : 0);
  NLS_TRY(ensure(ctx, ctx->part, (size_t)ntiles * cap * 8));
  OpVariance::Params vp;
  vp.A = psi_operand(ctx, g, rows);
  vp.B = basis_operand(g, bs.bt);
  vp.n_rows = rows;
  vp.m = full_cols;
  vp.bias_r = bs.bias_r;
  vp.bias_i = bs.bias_i;
  vp.w = bs.w;
  vp.part = (double*)ctx->part.p;
  vp.part_ld = cap;
  vp.b_upper = b_upper;
  NLS_TRY((launch_gemm<MODE_COMPLEX, OpVariance>(ctx, vp, dim3(gemm_tiles, (rows + BM - 1) / BM), rows, 2 * g.Np,
                                                  NLS_PROF_VARIANCE, "variance")));
  if (full_cols < g.m) {
    ProfScope scope(ctx, NLS_PROF_VARIANCE);
    project_tail_kernel<<<(rows * 32 + 255) / 256, 256, 0, ctx->stream>>>(
        (const double*)ctx->psi.p, 2LL * g.Dp, g.Dp, rows, g.D, bs.bt, g.Dp, g.Np, full_cols, g.m, bs.bias_r, bs.bias_i, 1,
        nullptr, nullptr, 0.0, nullptr, nullptr, 0, bs.w, (double*)ctx->part.p + (size_t)gemm_tiles * cap);
    NLS_TRY(check_launch(ctx, "project_tail_kernel"));
  }
  rowsum_reduce_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>((const double*)ctx->part.p, ntiles, cap, rows,
                                                                   sigma2_out);
  return check_launch(ctx, "rowsum_reduce_kernel");
}

extern "C" int nls_primal_finalize(nls_ctx* ctx, const double* X, const double* y, const double* s, int64_t n, int d,
                                   const double* shift, const double* W, int D, const double* Q, const double* lam,
                                   double inv_c, double gamma, const double* beta_eig, const double* beta,
                                   int is_classifier, const double* sigma2_in, double* loo_res_out,
                                   double* yhat_loo_out, double* leverage_out, double* resid_out,
                                   double* loo_std_out) {
  NLS_TRY(check_map_args(ctx, X, n, d, shift, W, D));
  if (!y || !s || !beta_eig || !beta || !loo_res_out || !yhat_loo_out || !leverage_out || !resid_out || !loo_std_out)
    return fail(NLS_ERR_INVALID, "null pointer");
  if (!sigma2_in && (!Q || !lam)) return fail(NLS_ERR_INVALID, "Q and lam are required when sigma2_in is null");
  const MapGeom g = geom(d, D);
  NLS_TRY(prep_weights(ctx, g, W));
  BasisScratch bs;
  if (!sigma2_in) {
    NLS_TRY(prep_basis(ctx, g, Q, &bs));
    variance_weights_kernel<<<(g.m + 255) / 256, 256, 0, ctx->stream>>>(lam, g.m, inv_c, gamma, bs.w);
    NLS_TRY(check_launch(ctx, "variance_weights_kernel"));
  }
  OzBasis oz{};
  const bool use_oz = !sigma2_in && oz_usable(ctx, g);
  if (use_oz) NLS_TRY(oz_prep_basis(ctx, g, bs, tail_split(g.m), &oz));
  const long long cap = ctx->chunk_rows;
  NLS_TRY(ensure(ctx, ctx->psi, (size_t)cap * 2 * g.Dp * 8));
  NLS_TRY(ensure(ctx, ctx->rowtmp, (size_t)3 * cap * 8));
  double* sigma2_tmp = (double*)ctx->rowtmp.p;
  double* num = sigma2_tmp + cap;
  double* fit = num + cap;
  // With the sigma^2 stash the only thing this pass needs from phi is Re(phi beta) for the two coefficient vectors:
  // they are accumulated in the feature-map epilogue and the chunk is neither written nor read back.
  RowDots dots{};
  if (sigma2_in) {
    dots.x[0] = beta_eig; dots.y[0] = beta_eig + 1; dots.ysign[0] = 1.0;
    dots.x[1] = beta;     dots.y[1] = beta + 1;     dots.ysign[1] = 1.0;
    dots.stride = 2; dots.count = 2; dots.ld = cap;
    NLS_TRY(rowdots_buffer(ctx, g, &dots.out));
  }
  for (int64_t i0 = 0; i0 < n; i0 += cap) {
    const int rows = (int)std::min<int64_t>(cap, n - i0);
    const double* sigma2 = sigma2_in ? sigma2_in + i0 : sigma2_tmp;
    if (sigma2_in) {
      NLS_TRY(feature_chunk(ctx, g, X + i0 * d, shift, rows, FM_NONE, nullptr, 0, 0, nullptr, &dots));
      rowdots_reduce_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(dots.out, (g.D + BN - 1) / BN, cap, rows,
                                                                        beta_eig + 2 * D, beta + 2 * D, num, fit);
      NLS_TRY(check_launch(ctx, "rowdots_reduce_kernel"));
    } else {
      NLS_TRY(feature_chunk(ctx, g, X + i0 * d, shift, rows, FM_PLANAR, (double*)ctx->psi.p, 2LL * g.Dp, g.Dp, nullptr));
      NLS_TRY(variance_chunk(ctx, g, bs, rows, sigma2_tmp, 0, use_oz ? &oz : nullptr));
      gemv_pair_kernel<<<(rows * 32 + 255) / 256, 256, 0, ctx->stream>>>((const double*)ctx->psi.p, 2LL * g.Dp, g.Dp,
                                                                        rows, D, beta_eig, beta, num, fit);
      NLS_TRY(check_launch(ctx, "gemv_pair_kernel"));
    }
    finalize_rows_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(rows, y + i0, s + i0, sigma2, num, fit,
                                                                     is_classifier, loo_res_out + i0,
                                                                     yhat_loo_out + i0, leverage_out + i0,
                                                                     resid_out + i0, loo_std_out + i0);
    NLS_TRY(check_launch(ctx, "finalize_rows_kernel"));
  }
  return NLS_OK;
}

extern "C" int nls_primal_predict(nls_ctx* ctx, const double* X, int64_t n, int d, const double* shift,
                                  const double* W, int D, const double* beta, const double* B, const double* w,
                                  int b_upper, double* yhat_out, double* sigma_out) {
  NLS_TRY(check_map_args(ctx, X, n, d, shift, W, D));
  if (yhat_out && !beta) return fail(NLS_ERR_INVALID, "beta is required for yhat_out");
  if (sigma_out && (!B || !w)) return fail(NLS_ERR_INVALID, "B and w are required for sigma_out");
  const MapGeom g = geom(d, D);
  NLS_TRY(prep_weights(ctx, g, W));
  BasisScratch bs;
  if (sigma_out) {
    NLS_TRY(prep_basis(ctx, g, B, &bs, b_upper));
    CUDA_TRY(cudaMemcpyAsync(bs.w, w, (size_t)g.m * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  OzBasis oz{};
  const bool use_oz = sigma_out && oz_usable(ctx, g);
  if (use_oz) NLS_TRY(oz_prep_basis(ctx, g, bs, tail_split(g.m), &oz));
  const long long cap = ctx->chunk_rows;
  NLS_TRY(ensure(ctx, ctx->psi, (size_t)cap * 2 * g.Dp * 8));
  NLS_TRY(ensure(ctx, ctx->rowtmp, (size_t)3 * cap * 8));
  double* sigma2 = (double*)ctx->rowtmp.p;
  // yhat = Re(phi beta) rides in the feature-map epilogue; without sigma_out the chunk is not even written.
  RowDots dots{};
  if (yhat_out) {
    dots.x[0] = beta; dots.y[0] = beta + 1; dots.ysign[0] = 1.0;
    dots.x[1] = beta; dots.y[1] = beta + 1; dots.ysign[1] = 1.0;
    dots.stride = 2; dots.count = 1; dots.ld = cap;
    NLS_TRY(rowdots_buffer(ctx, g, &dots.out));
  }
  for (int64_t i0 = 0; i0 < n; i0 += cap) {
    const int rows = (int)std::min<int64_t>(cap, n - i0);
    NLS_TRY(feature_chunk(ctx, g, X + i0 * d, shift, rows, sigma_out ? FM_PLANAR : FM_NONE, (double*)ctx->psi.p,
                          2LL * g.Dp, g.Dp, nullptr, yhat_out ? &dots : nullptr));
    if (yhat_out) {
      rowdots_reduce_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(dots.out, (g.D + BN - 1) / BN, cap, rows,
                                                                        beta + 2 * D, nullptr, yhat_out + i0, nullptr);
      NLS_TRY(check_launch(ctx, "rowdots_reduce_kernel"));
    }
    if (sigma_out) {
      NLS_TRY(variance_chunk(ctx, g, bs, rows, sigma2, b_upper, use_oz ? &oz : nullptr));
      sqrt_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(sigma2, rows, sigma_out + i0);
      NLS_TRY(check_launch(ctx, "sqrt_kernel"));
    }
  }
  return NLS_OK;
}

extern "C" int nls_quantile_epilogue(nls_ctx* ctx, const double* yhat, const double* sigma, int64_t n,
                                     const double* beta_abs, const double* beta_rel, const double* bias_abs,
                                     const double* bias_rel, int Q, int is_regressor, const double* iso_x,
                                     const double* iso_y, int n_iso, double* out) {
  if (!ctx || !yhat || !sigma || !beta_abs || !beta_rel || !bias_abs || !bias_rel || !out)
    return fail(NLS_ERR_INVALID, "null pointer");
  if (Q < 1) return fail(NLS_ERR_INVALID, "Q must be positive");
  if (!is_regressor && (!iso_x || !iso_y || n_iso < 1)) return fail(NLS_ERR_INVALID, "classifier needs isotonic thresholds");
  if (n < 1) return NLS_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  quantile_epilogue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      yhat, sigma, n, beta_abs, beta_rel, bias_abs, bias_rel, Q, is_regressor, iso_x, iso_y, n_iso, out);
  return check_launch(ctx, "quantile_epilogue_kernel");
}

// ---------------------------------------------------------------------------------------------
// Dual path
// ---------------------------------------------------------------------------------------------
static int rbf_block(nls_ctx* ctx, const double* xa, const double* na, int rows_a, const double* xb, const double* nb,
                     int rows_b, int p, int pp, double add, int force_diag, double* out, long long ld) {
  OpStore<true>::Params rp;
  rp.A = Operand{xa, pp, rows_a, p, 0, 0};
  rp.B = Operand{xb, pp, rows_b, p, 0, 0};
  rp.n_rows = rows_a;
  rp.n_cols = rows_b;
  rp.norm_a = na;
  rp.norm_b = nb;
  rp.add = add;
  rp.zero_diag_to = force_diag ? 0 : -1;
  rp.out = out;
  rp.ld = ld;
  return launch_gemm<MODE_REAL, OpStore<true>>(ctx, rp, dim3((rows_b + BN - 1) / BN, (rows_a + BM - 1) / BM), rows_a,
                                               rows_b, NLS_PROF_OTHER, "rbf");
}

static int pad_and_norm(nls_ctx* ctx, const double* X, long long rows, int p, int pp, double* xpad, double* norms) {
  pad_rows_kernel<<<grid_for(rows * pp), 256, 0, ctx->stream>>>(X, p, rows, p, pp, xpad);
  NLS_TRY(check_launch(ctx, "pad_rows_kernel"));
  row_sqnorm_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, ctx->stream>>>(xpad, pp, rows, p, norms);
  return check_launch(ctx, "row_sqnorm_kernel");
}

extern "C" int nls_dual_sweep(nls_ctx* ctx, const double* Xt, int n, int p, const double* y, const double* s,
                              const double* sn, const double* gammas, int G, int is_classifier, double* sums_out,
                              double* yhat_loo_out, double* lam_out) {
  if (!ctx || !Xt || !y || !s || !sn || !gammas || !sums_out || !yhat_loo_out || !lam_out)
    return fail(NLS_ERR_INVALID, "null pointer");
  if (n < 2 || p < 1 || G < 1) return fail(NLS_ERR_INVALID, "bad shape (n=%d p=%d G=%d)", n, p, G);
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->dual_n = 0;
  const long long ldn = round_up(n, 16);
  const int pp = (int)round_up(p, 16);
  const int Gp = (int)round_up(G, BN);
  const size_t nn = (size_t)n * ldn * 8;
  NLS_TRY(ensure(ctx, ctx->d_xpad, (size_t)n * pp * 8));
  NLS_TRY(ensure(ctx, ctx->d_norm, (size_t)2 * ldn * 8 + (size_t)65536 * 8));
  NLS_TRY(ensure(ctx, ctx->d_fm, 2 * nn));
  NLS_TRY(ensure(ctx, ctx->d_sq, nn));
  NLS_TRY(ensure(ctx, ctx->d_sqt, nn));
  NLS_TRY(ensure(ctx, ctx->d_g1, nn));
  NLS_TRY(ensure(ctx, ctx->d_ab, 2 * nn));
  NLS_TRY(ensure(ctx, ctx->d_ra, (size_t)(Gp + G) * ldn * 8));
  NLS_TRY(ensure(ctx, ctx->d_vec, (size_t)4 * ldn * 8));
  NLS_TRY(ensure(ctx, ctx->d_ng, (size_t)2 * n * G * 8));
  NLS_TRY(ensure(ctx, ctx->solver_mat, (size_t)n * n * 8 + 64));
  const int mtiles = (n + BM - 1) / BM;
  NLS_TRY(ensure(ctx, ctx->part, (size_t)mtiles * 3 * G * 8));
  double* xpad = (double*)ctx->d_xpad.p;
  double* norms = (double*)ctx->d_norm.p;
  double* F = (double*)ctx->d_fm.p;
  double* M = F + (size_t)n * ldn;
  double* SQ = (double*)ctx->d_sq.p;
  double* SQt = (double*)ctx->d_sqt.p;
  double* G1 = (double*)ctx->d_g1.p;
  double* AB = (double*)ctx->d_ab.p;
  double* alooT = (double*)ctx->d_ra.p;
  double* Rt = alooT + (size_t)Gp * ldn;
  double* qsy = (double*)ctx->d_vec.p;
  double* colsum = qsy + ldn;
  double* lam = colsum + ldn;
  double* aloo = (double*)ctx->d_ng.p;
  double* hd = aloo + (size_t)n * G;
  double* S = (double*)ctx->solver_mat.p;
  int* info = (int*)(S + (size_t)n * n);
  // F = rbf(Xt) + 1   (:261)
  NLS_TRY(pad_and_norm(ctx, Xt, n, p, pp, xpad, norms));
  NLS_TRY(rbf_block(ctx, xpad, norms, n, xpad, norms, n, p, pp, 1.0, 1, F, ldn));
  // lam, Q = eigh(sn F sn)   (:265): hand-written tridiagonalisation + divide and conquer (csrc/hetrd.cuh,
  // csrc/stedc.cuh); cuSOLVER's Dsyevd only as the comparator (NLS_EIG=cusolver / nls_ctx_set_eigensolver(1)).
  scale_sym_kernel<<<grid_for((long long)n * n), 256, 0, ctx->stream>>>(F, ldn, n, sn, S);
  NLS_TRY(check_launch(ctx, "scale_sym_kernel"));
  if (ctx->eig_kind == 1) {
    int lwork = 0;
    SOLVER_TRY(cusolverDnDsyevd_bufferSize(ctx->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, S, n, lam,
                                           &lwork));
    NLS_TRY(ensure(ctx, ctx->solver_ws, (size_t)lwork * 8));
    SOLVER_TRY(cusolverDnDsyevd(ctx->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, S, n, lam,
                                (double*)ctx->solver_ws.p, lwork, info));
    int h_info = 0;
    CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_info != 0) return fail(NLS_ERR_SOLVER, "symmetric eigensolver failed: info = %d", h_info);
    // Column-major eigenvectors = row-major Q^T; SQ^T[k, j] = Q[j, k] sn_j, SQ = its transpose.
    scale_cols_kernel<<<grid_for((long long)n * n), 256, 0, ctx->stream>>>(S, n, n, n, sn, ldn, SQt);
    NLS_TRY(check_launch(ctx, "scale_cols_kernel"));
  } else {
    double* Y = nullptr;  // row k = eigenvector k, i.e. row-major Q^T
    long long ldy = 0, yplane = 0;
    NLS_TRY(heev_tridiag<false>(ctx, S, n, n, 1.0, lam, &Y, &ldy, &yplane));
    scale_cols_kernel<<<grid_for((long long)n * n), 256, 0, ctx->stream>>>(Y, ldy, n, n, sn, ldn, SQt);
    NLS_TRY(check_launch(ctx, "scale_cols_kernel"));
  }
  CUDA_TRY(cudaMemcpyAsync(lam_out, lam, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  {
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, ctx->stream>>>(SQt, ldn, n, n, SQ, ldn);
    NLS_TRY(check_launch(ctx, "transpose_kernel"));
  }
  const unsigned warp_grid = (unsigned)(((long long)n * 32 + 255) / 256);
  gemv_rows_kernel<<<warp_grid, 256, 0, ctx->stream>>>(SQt, ldn, n, n, y, 0.0, qsy);   // Q^T (sn*y), :268
  NLS_TRY(check_launch(ctx, "gemv_rows_kernel"));
  gemv_rows_kernel<<<warp_grid, 256, 0, ctx->stream>>>(SQt, ldn, n, n, nullptr, 0.0, colsum);
  NLS_TRY(check_launch(ctx, "gemv_rows_kernel"));
  dual_ab_kernel<<<grid_for((long long)n * n), 256, 0, ctx->stream>>>(SQ, ldn, n, qsy, AB);
  NLS_TRY(check_launch(ctx, "dual_ab_kernel"));
  build_rt_kernel<<<grid_for((long long)G * ldn), 256, 0, ctx->stream>>>(gammas, lam, G, n, ldn, Rt);
  NLS_TRY(check_launch(ctx, "build_rt_kernel"));
  // alpha_loo = alpha_mat r (:285), hdiag = (SQ*SQ) r (:272-281)
  {
    OpDualCoef::Params cp;
    cp.A = Operand{AB, ldn, 2 * n, n, n, 0};
    cp.B = Operand{Rt, ldn, G, n, 0, 0};
    cp.n = n;
    cp.G = G;
    cp.aloo = aloo;
    cp.alooT = alooT;
    cp.ldT = ldn;
    cp.hd = hd;
    cp.eps = 2.220446049250313e-16;
    NLS_TRY((launch_gemm<MODE_DUAL_A, OpDualCoef>(ctx, cp, dim3((G + BN - 1) / BN, mtiles), 2 * n, G, NLS_PROF_SWEEP,
                                                   "dual_coef")));
  }
  // G1 = F0 SQ, M = G1 * SQ
  {
    OpDualCross::Params xp;
    xp.A = Operand{F, ldn, n, n, 0, 0};
    xp.B = Operand{SQt, ldn, n, n, 0, 0};
    xp.n = n;
    xp.SQ = SQ;
    xp.ld = ldn;
    xp.fdiag = 2.0;  // F_ii = exp(0) + 1
    xp.G1 = G1;
    xp.M = M;
    NLS_TRY((launch_gemm<MODE_REAL, OpDualCross>(ctx, xp, dim3((n + BN - 1) / BN, mtiles), n, n, NLS_PROF_PROJECT,
                                                  "dual_cross")));
  }
  // yhat_loo = -(M r)/hdiag * alpha_loo + F0 alpha_loo (:286) and the error sums (:287-302)
  {
    OpDualSweep::Params sp;
    sp.A = Operand{F, ldn, 2 * n, n, n, 0};
    sp.B = Operand{alooT, ldn, Gp + G, n, Gp, 0};
    sp.n_rows = n;
    sp.G = G;
    sp.y = y;
    sp.s = s;
    sp.is_classifier = is_classifier;
    sp.part = (double*)ctx->part.p;
    sp.aloo = aloo;
    sp.hd = hd;
    sp.fdiag = 2.0;
    sp.yhat_loo = yhat_loo_out;
    NLS_TRY((launch_gemm<MODE_PAIR, OpDualSweep>(ctx, sp, dim3((G + BN - 1) / BN, mtiles), 2 * n, Gp + G,
                                                  NLS_PROF_SWEEP, "dual_sweep")));
  }
  CUDA_TRY(cudaMemsetAsync(sums_out, 0, (size_t)3 * G * 8, ctx->stream));
  sweep_reduce_kernel<<<(3 * G + 255) / 256, 256, 0, ctx->stream>>>((const double*)ctx->part.p, mtiles, G, sums_out);
  NLS_TRY(check_launch(ctx, "sweep_reduce_kernel"));
  ctx->dual_n = n;
  ctx->dual_y = y;
  ctx->dual_sn = sn;
  return NLS_OK;
}

extern "C" int nls_dual_finalize(nls_ctx* ctx, int n, const double* y, const double* sn, double gamma,
                                 double* alpha_out, double* alpha_eig_out, double* U_out, double* Falpha_out,
                                 double* sigma2_out, double* Bt_out, double* w_out) {
  if (!ctx || !y || !sn || !alpha_out) return fail(NLS_ERR_INVALID, "null pointer");
  if (ctx->dual_n != n || n < 2) return fail(NLS_ERR_INVALID, "nls_dual_finalize must follow nls_dual_sweep with the same n");
  if (ctx->dual_y != (const void*)y || ctx->dual_sn != (const void*)sn)
    return fail(NLS_ERR_INVALID, "nls_dual_finalize: y / sn are not the arrays of the last nls_dual_sweep on this context "
                                 "(the sweep's kernel matrix and eigenbasis are context state: finish one model before starting another)");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const long long ldn = round_up(n, 16);
  double* F = (double*)ctx->d_fm.p;
  double* SQ = (double*)ctx->d_sq.p;
  double* SQt = (double*)ctx->d_sqt.p;
  double* G1 = (double*)ctx->d_g1.p;
  double* qsy = (double*)ctx->d_vec.p;
  double* colsum = qsy + ldn;
  double* lam = colsum + ldn;
  double* xk = lam + ldn;
  const unsigned warp_grid = (unsigned)(((long long)n * 32 + 255) / 256);
  // alpha from the eigen-expansion (:311): alpha = SQ (qsy / (gamma + lam))
  dual_coef_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(qsy, lam, n, gamma, xk);
  NLS_TRY(check_launch(ctx, "dual_coef_kernel"));
  double* aeig = alpha_eig_out ? alpha_eig_out : alpha_out;
  gemv_rows_kernel<<<warp_grid, 256, 0, ctx->stream>>>(SQ, ldn, n, n, xk, 0.0, aeig);
  NLS_TRY(check_launch(ctx, "gemv_rows_kernel"));
  if (U_out) {
    // Re-solve with a Cholesky factorisation of gamma S^-2 + F (:313-314); U_out is cho_factor's layout.
    pad_rows_kernel<<<grid_for((long long)n * n), 256, 0, ctx->stream>>>(F, ldn, n, n, n, U_out);
    NLS_TRY(check_launch(ctx, "pad_rows_kernel"));
    dual_add_diag_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(U_out, n, sn, gamma);
    NLS_TRY(check_launch(ctx, "dual_add_diag_kernel"));
    if (ctx->eig_kind != 1) {  // hand-written blocked Cholesky + triangular solves (csrc/potrf.cuh)
      NLS_TRY(chol_factor<false>(ctx, U_out, n, n));
      CUDA_TRY(cudaMemcpyAsync(alpha_out, y, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      NLS_TRY(chol_solve<false>(ctx, U_out, n, n, alpha_out));
    } else {  // cuSOLVER comparator
      int lwork = 0;
      SOLVER_TRY(cusolverDnDpotrf_bufferSize(ctx->solver, CUBLAS_FILL_MODE_LOWER, n, U_out, n, &lwork));
      NLS_TRY(ensure(ctx, ctx->solver_ws, (size_t)lwork * 8 + 64));
      int* info = (int*)((char*)ctx->solver_ws.p + (size_t)lwork * 8);
      SOLVER_TRY(cusolverDnDpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, n, U_out, n, (double*)ctx->solver_ws.p, lwork, info));
      int h_info = 0;
      CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      if (h_info != 0) return fail(NLS_ERR_SOLVER, "dual Cholesky factorisation failed: info = %d", h_info);
      CUDA_TRY(cudaMemcpyAsync(alpha_out, y, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      SOLVER_TRY(cusolverDnDpotrs(ctx->solver, CUBLAS_FILL_MODE_LOWER, n, 1, U_out, n, alpha_out, n, info));
    }
  } else if (alpha_eig_out) {
    CUDA_TRY(cudaMemcpyAsync(alpha_out, alpha_eig_out, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (Falpha_out) {  // F alpha (:315)
    gemv_rows_kernel<<<warp_grid, 256, 0, ctx->stream>>>(F, ldn, n, n, alpha_out, 0.0, Falpha_out);
    NLS_TRY(check_launch(ctx, "gemv_rows_kernel"));
  }
  if (sigma2_out) {  // :321-322 in the eigenbasis
    dual_sigma2_kernel<<<warp_grid, 256, 0, ctx->stream>>>(G1, SQ, ldn, n, colsum, lam, 2.0, gamma, sigma2_out);
    NLS_TRY(check_launch(ctx, "dual_sigma2_kernel"));
  }
  if (Bt_out) {
    pad_rows_kernel<<<grid_for((long long)n * n), 256, 0, ctx->stream>>>(SQt, ldn, n, n, n, Bt_out);
    NLS_TRY(check_launch(ctx, "pad_rows_kernel"));
  }
  if (w_out) {
    variance_weights_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(lam, n, 1.0, gamma, w_out);
    NLS_TRY(check_launch(ctx, "variance_weights_kernel"));
  }
  return NLS_OK;
}

extern "C" int nls_dual_predict(nls_ctx* ctx, const double* Xq, int64_t nq, const double* Xt, int n, int p,
                                const double* alpha, double alpha_sum, const double* Bt, const double* w,
                                double* yhat_out, double* sigma_out) {
  if (!ctx || !Xq || !Xt) return fail(NLS_ERR_INVALID, "null pointer");
  if (yhat_out && !alpha) return fail(NLS_ERR_INVALID, "alpha is required for yhat_out");
  if (sigma_out && (!Bt || !w)) return fail(NLS_ERR_INVALID, "Bt and w are required for sigma_out");
  if (nq < 1 || n < 1 || p < 1) return fail(NLS_ERR_INVALID, "bad shape");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const long long ldn = round_up(n, 16);
  const int pp = (int)round_up(p, 16);
  const int cq = (int)std::min<int64_t>(std::min<int64_t>(ctx->chunk_rows, 8192), round_up(nq, 128));
  const int ntiles = (n + BN - 1) / BN;
  NLS_TRY(ensure(ctx, ctx->d_xpad, (size_t)n * pp * 8));
  NLS_TRY(ensure(ctx, ctx->d_xq, (size_t)cq * pp * 8));
  NLS_TRY(ensure(ctx, ctx->d_norm, (size_t)2 * ldn * 8 + (size_t)65536 * 8));
  NLS_TRY(ensure(ctx, ctx->d_kq, (size_t)cq * ldn * 8));
  NLS_TRY(ensure(ctx, ctx->rowtmp, (size_t)3 * std::max<long long>(cq, ctx->chunk_rows) * 8));
  ctx->dual_n = 0;  // d_xpad / d_norm are shared with the fit state
  double* xt = (double*)ctx->d_xpad.p;
  double* nt = (double*)ctx->d_norm.p;
  double* nqv = nt + 2 * ldn;
  double* xq = (double*)ctx->d_xq.p;
  double* Kq = (double*)ctx->d_kq.p;
  double* q = (double*)ctx->rowtmp.p;
  NLS_TRY(pad_and_norm(ctx, Xt, n, p, pp, xt, nt));
  double* btp = nullptr;
  if (sigma_out) {
    NLS_TRY(ensure(ctx, ctx->d_btp, (size_t)n * ldn * 8));
    NLS_TRY(ensure(ctx, ctx->part, (size_t)ntiles * cq * 8));
    btp = (double*)ctx->d_btp.p;
    pad_rows_kernel<<<grid_for((long long)n * ldn), 256, 0, ctx->stream>>>(Bt, n, n, n, (int)ldn, btp);
    NLS_TRY(check_launch(ctx, "pad_rows_kernel"));
  }
  for (int64_t i0 = 0; i0 < nq; i0 += cq) {
    const int rows = (int)std::min<int64_t>(cq, nq - i0);
    NLS_TRY(pad_and_norm(ctx, Xq + i0 * p, rows, p, pp, xq, nqv));
    NLS_TRY(rbf_block(ctx, xq, nqv, rows, xt, nt, n, p, pp, 0.0, 0, Kq, ldn));  // :474 / :669
    const unsigned warp_grid = (unsigned)(((long long)rows * 32 + 255) / 256);
    if (yhat_out) {  // K alpha + sum(alpha), :670-671
      gemv_rows_kernel<<<warp_grid, 256, 0, ctx->stream>>>(Kq, ldn, rows, n, alpha, alpha_sum, yhat_out + i0);
      NLS_TRY(check_launch(ctx, "gemv_rows_kernel"));
    }
    if (sigma_out) {  // sqrt(1 - sum_k (K B^T)_ik^2 w_k), :475-477
      OpRowQuad::Params vp;
      vp.A = Operand{Kq, ldn, rows, n, 0, 0};
      vp.B = Operand{btp, ldn, n, n, 0, 0};
      vp.n_rows = rows;
      vp.m = n;
      vp.bias_r = nullptr;
      vp.bias_i = nullptr;
      vp.w = w;
      vp.part = (double*)ctx->part.p;
      vp.part_ld = cq;
      vp.b_upper = 0;
      NLS_TRY((launch_gemm<MODE_REAL, OpRowQuad>(ctx, vp, dim3(ntiles, (rows + BM - 1) / BM), rows, n, NLS_PROF_VARIANCE,
                                                  "dual_variance")));
      rowsum_reduce_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>((const double*)ctx->part.p, ntiles, cq, rows, q);
      NLS_TRY(check_launch(ctx, "rowsum_reduce_kernel"));
      one_minus_sqrt_kernel<<<(rows + 255) / 256, 256, 0, ctx->stream>>>(q, rows, sigma_out + i0);
      NLS_TRY(check_launch(ctx, "one_minus_sqrt_kernel"));
    }
  }
  return NLS_OK;
}

// ---------------------------------------------------------------------------------------------
// Supervised affine pre-pass: per-bin weighted median / MAD of every feature
// ---------------------------------------------------------------------------------------------
extern "C" int nls_bin_median_stats(nls_ctx* ctx, const double* X, int64_t n, int d, const int64_t* perm,
                                    const double* w, const int* tiles, int ntiles, const int* bin_tiles, int nbins,
                                    const double* thresh, double* stats_out, double* wtot_out) {
  if (!ctx || !X || !perm || !w || !tiles || !bin_tiles || !stats_out || !wtot_out)
    return fail(NLS_ERR_INVALID, "null pointer");
  if (n < 1 || d < 1 || ntiles < 1 || nbins < 1) return fail(NLS_ERR_INVALID, "bad shape");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const long long nd = (long long)nbins * d;
  NLS_TRY(ensure(ctx, ctx->bs_part, (size_t)ntiles * 7 * d * 8));
  NLS_TRY(ensure(ctx, ctx->bs_keys, (size_t)3 * nd * 8));
  double* partial = (double*)ctx->bs_part.p;
  unsigned long long* lo = (unsigned long long*)ctx->bs_keys.p;
  unsigned long long* hi = lo + nd;
  unsigned long long* mid = hi + nd;
  const dim3 grid(ntiles, (d + BS_COLS - 1) / BS_COLS);
  const int sgrid = (int)((nd + 255) / 256);
  const BsTile* tl = (const BsTile*)tiles;
  const int2* bt = (const int2*)bin_tiles;
  ProfScope scope(ctx, NLS_PROF_OTHER);
  const long long* pm = (const long long*)perm;
  auto count = [&]() {  // columns per lane: 1, 2, 4 or 8
    if (d <= 32) bs_count_kernel<1><<<dim3(ntiles, 1), 256, 0, ctx->stream>>>(X, d, pm, w, tl, mid, partial);
    else if (d <= 64) bs_count_kernel<2><<<dim3(ntiles, 1), 256, 0, ctx->stream>>>(X, d, pm, w, tl, mid, partial);
    else if (d <= 128) bs_count_kernel<4><<<dim3(ntiles, 1), 256, 0, ctx->stream>>>(X, d, pm, w, tl, mid, partial);
    else bs_count_kernel<8><<<dim3(ntiles, (d + 255) / 256), 256, 0, ctx->stream>>>(X, d, pm, w, tl, mid, partial);
  };
  // Total weight per (bin, column): every key is <= the all-ones key.
  CUDA_TRY(cudaMemsetAsync(mid, 0xff, (size_t)nd * 8, ctx->stream));
  count();
  bs_step_kernel<<<sgrid, 256, 0, ctx->stream>>>(partial, bt, nbins, d, -1, thresh, wtot_out, lo, hi, mid);
  ctx->launches += 2;
  // 64 bisection steps on the 64-bit key: afterwards lo == hi == key of the crossing value v*.
  for (int it = 0; it < 64; ++it) {
    count();
    bs_step_kernel<<<sgrid, 256, 0, ctx->stream>>>(partial, bt, nbins, d, it, thresh, wtot_out, lo, hi, mid);
    ctx->launches += 2;
  }
  bs_stats_kernel<<<grid, 256, 0, ctx->stream>>>(X, d, (const long long*)perm, w, tl, lo, partial);
  bs_stats_reduce_kernel<<<sgrid, 256, 0, ctx->stream>>>(partial, bt, nbins, d, lo, stats_out);
  ctx->launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(NLS_ERR_CUDA, "bin statistics launch failed: %s", cudaGetErrorString(e));
  return NLS_OK;
}

extern "C" int nls_bin_mad(nls_ctx* ctx, const double* X, int64_t n, int d, const int64_t* perm, const double* w,
                           const int* tiles, int ntiles, const int* bin_tiles, int nbins, const double* centre,
                           double* spread_out) {
  if (!ctx || !X || !perm || !w || !tiles || !bin_tiles || !centre || !spread_out)
    return fail(NLS_ERR_INVALID, "null pointer");
  if (n < 1 || d < 1 || ntiles < 1 || nbins < 1) return fail(NLS_ERR_INVALID, "bad shape");
  CUDA_TRY(cudaSetDevice(ctx->device));
  NLS_TRY(ensure(ctx, ctx->bs_part, (size_t)ntiles * 7 * d * 8));
  double* partial = (double*)ctx->bs_part.p;
  const dim3 grid(ntiles, (d + BS_COLS - 1) / BS_COLS);
  ProfScope scope(ctx, NLS_PROF_OTHER);
  bs_mad_kernel<<<grid, 256, 0, ctx->stream>>>(X, d, (const long long*)perm, w, (const BsTile*)tiles, centre, partial);
  NLS_TRY(check_launch(ctx, "bs_mad_kernel"));
  bs_sum_kernel<<<(nbins * d + 255) / 256, 256, 0, ctx->stream>>>(partial, (const int2*)bin_tiles, nbins, d, spread_out);
  return check_launch(ctx, "bs_sum_kernel");
}

// ---------------------------------------------------------------------------------------------
// Micro-benchmark: resident-tile tcgen05.mma kind::i8 loop (INT8 tensor peak of this device), at the hardware's
// widest single-CTA shape (n_cols = 256) or at the Ozaki tile's shape (n_cols = 64).
// ---------------------------------------------------------------------------------------------
template <int N>
static int i8_peak(nls_ctx* ctx, int iters, double* tops_out) {
  const int smem = 1024 + (oz::TM + N) * 128;
  CUDA_TRY(cudaFuncSetAttribute(oz::i8_peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  oz::i8_peak_kernel<N><<<ctx->sm_count, 128, smem, ctx->stream>>>(iters / 8 + 1);
  float ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(a, ctx->stream));
    oz::i8_peak_kernel<N><<<ctx->sm_count, 128, smem, ctx->stream>>>(iters);
    CUDA_TRY(cudaEventRecord(b, ctx->stream));
    NLS_TRY(check_launch(ctx, "oz::i8_peak_kernel"));
    CUDA_TRY(cudaEventSynchronize(b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, a, b));
    ms = t < ms ? t : ms;
  }
  *tops_out = (double)ctx->sm_count * iters * 4.0 * (2.0 * oz::TM * N * 32) / (ms * 1e-3) / 1e12;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return NLS_OK;
}

extern "C" int nls_bench_i8_peak(nls_ctx* ctx, int iters, int n_cols, double* tops_out) {
  if (!ctx || !tops_out || iters < 1 || (n_cols != 64 && n_cols != 256)) return fail(NLS_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  return n_cols == 64 ? i8_peak<64>(ctx, iters, tops_out) : i8_peak<256>(ctx, iters, tops_out);
}

// ---------------------------------------------------------------------------------------------
// Micro-benchmark: register-resident DMMA loop (FP64 tensor peak of this device).
// ---------------------------------------------------------------------------------------------
extern "C" int nls_bench_dmma_peak(nls_ctx* ctx, int iters, double* tflops_out) {
  if (!ctx || !tflops_out || iters < 1) return fail(NLS_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  NLS_TRY(ensure(ctx, ctx->small, 4096));
  // Two resident CTAs of 8 warps per SM (4 warps per scheduler, 16 independent accumulator tiles each): one
  // full wave, no tail.
  const char* bps = getenv("NLS_PEAK_BLOCKS_PER_SM");
  const int blocks = ctx->sm_count * (bps && atoi(bps) > 0 ? atoi(bps) : 2), threads = 256;
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  // A short untimed launch ramps the clocks from idle; the first timed launch then sees the burst
  // state, the following ones the state under sustained load.  The best of all is reported.
  dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(iters / 8 + 1, (double*)ctx->small.p);
  float ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(a, ctx->stream));
    dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, (double*)ctx->small.p);
    CUDA_TRY(cudaEventRecord(b, ctx->stream));
    NLS_TRY(check_launch(ctx, "dmma_peak_kernel"));
    CUDA_TRY(cudaEventSynchronize(b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, a, b));
    ms = t < ms ? t : ms;
  }
  const double flops = (double)blocks * (threads / 32) * (double)iters * 16.0 * (2.0 * 8 * 8 * 4);
  *tflops_out = flops / (ms * 1e-3) / 1e12;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return NLS_OK;
}
