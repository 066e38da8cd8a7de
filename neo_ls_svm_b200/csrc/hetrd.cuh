// hetrd.cuh — Householder tridiagonalisation of a dense Hermitian / real symmetric matrix and the matching
// back-transformation: first and last stage of the hand-written eigensolver that replaces LAPACK behind
// scipy.linalg.eigh (_neo_ls_svm.py:120, complex m x m) and np.linalg.eigh (:265, real n x n).
//
// Blocked LAPACK-style reduction (zhetrd/zlatrd, lower form, full storage), restructured for one persistent
// cooperative kernel per panel of NB = 32 columns:
//   * per column ONE exchange of the raw updated column x and ONE grid-wide reduction of
//       [x^H x, x^H (A x), x^H a_1, V^H x, W^H x];
//     the Householder scalars (beta, tau) and everything that depends on them (v, V^H v, W^H v, v^H A v, w) follow
//     from those sums ("deferred alpha": v = sigma (x - beta e_1) is linear in x), so a column costs two grid
//     barriers instead of the five dependent reductions of the textbook loop;
//   * the big operand, y_raw = A x over the trailing matrix, is row-slab distributed over all SMs (warp per row
//     segment, x staged in shared memory, fixed-order partial sums => bitwise reproducible);
//   * after the panel the trailing matrix takes the rank-2NB update A -= V W^H + W V^H as ONE GEMM on the FP64 DMMA
//     core (OpHer2k, K = 2 NB), and the reflectors are kept in the two K-contiguous layouts the back-transformation
//     GEMMs want (VT: one reflector per row; VR: one coordinate per row, conjugated).
// Back-transformation X = Q Z with Q = prod_p (I - V_p T_p V_p^H), on Y = X^T (row k = eigenvector k):
//   C1 = Y conj(V_p) (split-K GEMM, OpBT1)  ->  C2 = C1 T_p^T (small kernel)  ->  Y -= C2 V_p^T (GEMM, OpBT2).
#pragma once
#include <cuda_runtime.h>

#include "gemm_core.cuh"

namespace nls {

constexpr int HNB = 32;            // panel width
constexpr int HETRD_THREADS = 512;
constexpr int HETRD_WARPS = HETRD_THREADS / 32;
constexpr int HETRD_NSUM = 3 + 2 * HNB;  // grid-reduced sums per column
constexpr int HETRD_SPAN = 512;          // elements of one unrolled hemv step per warp (real: 32 lanes x 2 x 8)

// ---- scalar helpers: T = double (real symmetric) or double2 (complex Hermitian, interleaved re/im) -------------
template <bool C> struct HS;
template <> struct HS<false> {
  using T = double;
  static __device__ __forceinline__ T zero() { return 0.0; }
  static __device__ __forceinline__ T make(double re, double) { return re; }
  static __device__ __forceinline__ double re(T a) { return a; }
  static __device__ __forceinline__ double im(T) { return 0.0; }
  static __device__ __forceinline__ T conj(T a) { return a; }
  static __device__ __forceinline__ T add(T a, T b) { return a + b; }
  static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
  static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
  static __device__ __forceinline__ T cmul(T a, T b) { return a * b; }             // conj(a) * b
  static __device__ __forceinline__ T fma(T a, T b, T c) { return ::fma(a, b, c); }  // a * b + c
  static __device__ __forceinline__ T cfma(T a, T b, T c) { return ::fma(a, b, c); }  // conj(a) * b + c
  static __device__ __forceinline__ T scale(double s, T a) { return s * a; }
  static __device__ __forceinline__ double abs2(T a) { return a * a; }
  static __device__ __forceinline__ T ldcg(const T* p) { return __ldcg(p); }
  static __device__ __forceinline__ T shfl_xor(T a, int off) { return __shfl_xor_sync(0xffffffffu, a, off); }
};
template <> struct HS<true> {
  using T = double2;
  static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ T make(double re, double im) { return make_double2(re, im); }
  static __device__ __forceinline__ double re(T a) { return a.x; }
  static __device__ __forceinline__ double im(T a) { return a.y; }
  static __device__ __forceinline__ T conj(T a) { return make_double2(a.x, -a.y); }
  static __device__ __forceinline__ T add(T a, T b) { return make_double2(a.x + b.x, a.y + b.y); }
  static __device__ __forceinline__ T sub(T a, T b) { return make_double2(a.x - b.x, a.y - b.y); }
  static __device__ __forceinline__ T mul(T a, T b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
  static __device__ __forceinline__ T cmul(T a, T b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
  static __device__ __forceinline__ T fma(T a, T b, T c) {
    return make_double2(::fma(a.x, b.x, ::fma(-a.y, b.y, c.x)), ::fma(a.x, b.y, ::fma(a.y, b.x, c.y)));
  }
  static __device__ __forceinline__ T cfma(T a, T b, T c) {
    return make_double2(::fma(a.x, b.x, ::fma(a.y, b.y, c.x)), ::fma(a.x, b.y, ::fma(-a.y, b.x, c.y)));
  }
  static __device__ __forceinline__ T scale(double s, T a) { return make_double2(s * a.x, s * a.y); }
  static __device__ __forceinline__ double abs2(T a) { return a.x * a.x + a.y * a.y; }
  static __device__ __forceinline__ T ldcg(const T* p) { return __ldcg(p); }
  static __device__ __forceinline__ T shfl_xor(T a, int off) {
    return make_double2(__shfl_xor_sync(0xffffffffu, a.x, off), __shfl_xor_sync(0xffffffffu, a.y, off));
  }
};

template <bool C>
struct HetrdArgs {
  using T = typename HS<C>::T;
  T* A;             // n x n row-major work matrix (full storage), pitch lda elements
  long long lda;
  int n, k0, jb;    // panel = columns [k0, k0 + jb)
  T* V;             // n x HNB panel reflectors (row-major), rows > k0 used
  T* W;             // n x HNB
  double* d;        // n
  double* e;        // n - 1
  T* tau;           // n - 1
  T* Tfac;          // [panels][HNB][HNB] compact-WY factors (upper triangular)
  // planar GEMM operands (plane 1 = imaginary part, complex only), row pitch 2 HNB / ldv doubles
  double* PW;       // [planes][npad][2 HNB]: [V | W]
  double* RW;       // [planes][npad][2 HNB]: [W | V]
  double* VT;       // [planes][npad][ldv]: VT[c][k] = v_c[k]
  double* VR;       // [planes][npad][ldv]: VR[k][c] = conj(v_c[k])
  long long ldv;
  long long npad;
  // exchange buffers
  T* xbuf;          // n
  T* ybuf;          // n
  T* part;          // [HETRD_NSUM][gpad]
  int gpad;
  T* zpart;         // SYM only: [G][npad] column partials z_b[k] = sum_{rows r > k of CTA b} conj(A[r][k]) x[r]
  unsigned* bar;    // grid barrier counter (zeroed before every launch)
  long long* dbg;   // optional [8] cycle counters of CTA 0 (diagnostics), may be null
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void hetrd_grid_barrier(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned spins = 0;
    while (ld_acquire_u32(ctr) < target) {
      if (++spins > (1u << 30)) __trap();  // a lost CTA must not hang the GPU
    }
    __threadfence();
  }
  __syncthreads();
}

// Shared-memory layout of the panel kernel (dynamic): x staging area first (n elements of T), then fixed scratch.
template <bool C, bool SYM>
struct HetrdSmem {
  using T = typename HS<C>::T;
  T ysm[SYM ? HETRD_WARPS * 128 : 1];  // SYM: per-warp partial row sums of the strictly lower triangle
  T sums[HETRD_NSUM];            // grid-reduced: 0 xx, 1 xy, 2 xa, 3.. V^H x, 3+HNB.. W^H x
  T Vv[HNB], Wv[HNB];            // V^H v, W^H v
  T rowV[HNB], rowW[HNB];        // row c+1 of the panel matrices (becomes row c of the next column)
  T wred[HETRD_WARPS][2 * HNB];  // per-warp partials of V^H x | W^H x
  T psum[HETRD_THREADS];         // per-task partial row sums of A x
  T rowred[3][128];              // per-row contributions to xx, xy, xa
  T Tf[HNB][HNB + 1];            // compact-WY factor (CTA 0)
  double beta, sigma_re, sigma_im, tau_re, tau_im, alpha2_re, alpha2_im;
};

// Bytes of the x staging area: n elements rounded up to whole unrolled spans (+ one span of slack for the even base).
constexpr int HETRD_XPAD = 2048;  // x is zero-padded to whole column panels of the symmetric path (>= HETRD_SPAN)
template <bool C>
__host__ __device__ inline size_t hetrd_xs_bytes(int n) {
  return ((size_t)(n + 2 + HETRD_XPAD - 1) / HETRD_XPAD * HETRD_XPAD) * sizeof(typename HS<C>::T);
}

// SYM = false: every CTA reads full rows of the trailing matrix (latency-optimised, small matrices).
// SYM = true : only the strictly lower triangle is read (half the HBM traffic): element (r, k), k < r, contributes
//              A[r][k] x[k] to y[r] and conj(A[r][k]) x[r] to y[k]; the second kind is accumulated per CTA in
//              registers (a warp owns a fixed set of columns), written as one partial vector per CTA and summed over
//              the CTAs by the consumer of y[k].  Rows are dealt to the CTAs in two blocks (b and 2G-1-b) so that
//              every CTA gets the same share of the triangle.
template <bool C, bool SYM>
__global__ void __launch_bounds__(HETRD_THREADS, 1) hetrd_panel_kernel(const HetrdArgs<C> a) {
  using H = HS<C>;
  using T = typename H::T;
  extern __shared__ __align__(16) unsigned char hsm_raw[];
  T* xs = reinterpret_cast<T*>(hsm_raw);
  HetrdSmem<C, SYM>& sm = *reinterpret_cast<HetrdSmem<C, SYM>*>(hsm_raw + hetrd_xs_bytes<C>(a.n));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, b = blockIdx.x;
  const int n = a.n;
  const long long gtid = (long long)b * HETRD_THREADS + tid, gthreads = (long long)G * HETRD_THREADS;
  unsigned bar_target = 0;
  const int planes = C ? 2 : 1;
  (void)planes;

  if (b == 0)
    for (int t = tid; t < HNB * (HNB + 1); t += HETRD_THREADS) (&sm.Tf[0][0])[t] = H::zero();

  // ---- phase A of the panel's first column: x = A[k0+1:, k0] (the trailing matrix is up to date) ----
  for (long long r = a.k0 + gtid; r < n; r += gthreads) {
    const T v = a.A[r * a.lda + a.k0];
    if (r == a.k0) a.d[a.k0] = H::re(v);
    else a.xbuf[r] = v;
  }

  for (int j = 0; j < a.jb; ++j) {
    const int c = a.k0 + j;      // column being reduced
    const int f = c + 1;         // first row of x / of the trailing matrix
    const int n1 = n - f;        // trailing size
    long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0, tc4 = 0, tc5 = 0;
    const bool prof = a.dbg != nullptr && b == 0 && tid == 0;
    if (prof) tc0 = clock64();
    hetrd_grid_barrier(a.bar, bar_target);  // B1: x complete
    if (prof) tc1 = clock64();

    // ---- phase B: y_raw = A[f:, f:] x for this CTA's row slab, plus the partial sums --------------------
    const int kbase = f & ~1;    // x staged from an even index (vector loads in the real case)
    const int nk = n - kbase;
    constexpr int XPAD = SYM ? HETRD_XPAD : HETRD_SPAN;
    const int nk_pad = (nk + XPAD - 1) / XPAD * XPAD;  // zero tail: the unrolled loops need no guards
    for (int k = tid; k < nk_pad; k += HETRD_THREADS) {
      const int kk = kbase + k;
      xs[k] = (kk >= f && kk < n) ? H::ldcg(a.xbuf + kk) : H::zero();
    }
    if (SYM)
      for (int t = tid; t < HETRD_WARPS * 128; t += HETRD_THREADS) sm.ysm[t] = H::zero();
    __syncthreads();
    // Rows of this CTA: one contiguous slab, or (SYM) the two blocks b and 2G-1-b of 2G equal blocks.
    const int R = (n1 + G - 1) / G;
    int r_lo, rows_a, r_hi, rows_b;
    if (SYM) {  // block beta = rows [f + beta n1 / 2G, f + (beta + 1) n1 / 2G): heights differ by at most one
      const long long twoG = 2LL * G;
      r_lo = f + (int)((long long)b * n1 / twoG);
      rows_a = f + (int)((long long)(b + 1) * n1 / twoG) - r_lo;
      r_hi = f + (int)((twoG - 1 - b) * n1 / twoG);
      rows_b = f + (int)((twoG - b) * n1 / twoG) - r_hi;
    } else {
      r_lo = f + b * R;
      rows_a = max(0, min(n, r_lo + R) - r_lo);
      r_hi = 0;
      rows_b = 0;
    }
    const int rows = rows_a + rows_b;
    auto row_of = [&](int li) { return li < rows_a ? r_lo + li : r_hi + (li - rows_a); };
    int S = 1;
    if (!SYM) {
      // k segments per row: with few rows per CTA a row is spread over several warps (two tasks per warp is the aim);
      // a segment is a whole number of unrolled spans.
      S = 1;
      if (R < 2 * HETRD_WARPS) S = max(1, min(min(2 * HETRD_WARPS / max(R, 1), nk_pad / HETRD_SPAN), 16));
      const int seglen = ((nk_pad / HETRD_SPAN + S - 1) / S) * HETRD_SPAN;
      for (int task = warp; task < rows * S; task += HETRD_WARPS) {
        const int row = r_lo + task / S, seg = task % S;
        const int ka = seg * seglen, kb = min(nk_pad, ka + seglen);  // relative to kbase
        const T* arow = a.A + (long long)row * a.lda + kbase;
        T acc;
        if (C) {
          // HETRD_UNR independent 16-byte loads in flight per lane; indices past the row are clamped (x is zero there)
          const int kmax = (int)a.lda - 1 - kbase;
          T c0 = H::zero(), c1 = H::zero(), c2 = H::zero(), c3 = H::zero();
          for (int k = ka + lane; k < kb; k += 32 * 8) {
            T av[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = arow[min(k + 32 * u, kmax)];
#pragma unroll
            for (int u = 0; u < 8; u += 4) {
              c0 = H::fma(av[u], xs[k + 32 * u], c0);
              c1 = H::fma(av[u + 1], xs[k + 32 * (u + 1)], c1);
              c2 = H::fma(av[u + 2], xs[k + 32 * (u + 2)], c2);
              c3 = H::fma(av[u + 3], xs[k + 32 * (u + 3)], c3);
            }
          }
          acc = H::add(H::add(c0, c1), H::add(c2, c3));
        } else {
          const double* ar = reinterpret_cast<const double*>(arow);
          const double* xr = reinterpret_cast<const double*>(xs);
          const int kmax = (int)a.lda - 2 - kbase;  // even
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          for (int k = ka + 2 * lane; k < kb; k += 64 * 8) {  // kbase, ka even: 16-byte aligned pairs
            double2 av[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = *reinterpret_cast<const double2*>(ar + min(k + 64 * u, kmax));
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const double2 x0 = *reinterpret_cast<const double2*>(xr + k + 64 * u);
              const double2 x1 = *reinterpret_cast<const double2*>(xr + k + 64 * (u + 1));
              s0 = ::fma(av[u].x, x0.x, s0);
              s1 = ::fma(av[u].y, x0.y, s1);
              s2 = ::fma(av[u + 1].x, x1.x, s2);
              s3 = ::fma(av[u + 1].y, x1.y, s3);
            }
          }
          acc = H::make((s0 + s1) + (s2 + s3), 0.0);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc = H::add(acc, H::shfl_xor(acc, off));
        if (lane == 0) sm.psum[task] = acc;
      }
    } else {
      constexpr int VEC = C ? 1 : 2;                 // elements per 16-byte load
      constexpr int U = 2;                           // loads per row and tile
      constexpr int WC = 32 * VEC * U;               // columns of one warp in a column panel
      constexpr int CPW = WC * HETRD_WARPS;          // columns of a column panel (divides HETRD_XPAD)
      const int kmax = (int)a.lda - VEC - kbase;     // clamp for loads past the row (x is zero there)
      const double* xsd = reinterpret_cast<const double*>(xs);
      for (int cp0 = 0; cp0 < nk_pad; cp0 += CPW) {
        const int span0 = kbase + cp0 + warp * WC;    // first column of this warp's span (absolute)
        const int kw = cp0 + warp * WC + lane * VEC;  // this lane's first column, relative to kbase
        T zacc[U * VEC];
#pragma unroll
        for (int q = 0; q < U * VEC; ++q) zacc[q] = H::zero();
        if (span0 < n) {
          // x at this lane's columns: loaded once per span, shared by all rows
          double2 xv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) xv[u] = *reinterpret_cast<const double2*>(xsd + (C ? 2 : 1) * (kw + 32 * VEC * u));
          const int off0 = (C ? 2 : 1) * min(kw, kmax), off1 = (C ? 2 : 1) * min(kw + 32 * VEC, kmax);
#pragma unroll 1
          for (int grp = 0; grp < 2; ++grp) {
            const int g_lo = grp ? r_hi : r_lo, g_rows = grp ? rows_b : rows_a, li0 = grp ? rows_a : 0;
#pragma unroll 1
            for (int rb = 0; rb < g_rows; rb += 8) {
              const int r0 = g_lo + rb, nr = min(8, g_rows - rb);
              if (span0 >= r0 + nr - 1) continue;  // the span lies on or above the diagonal for these rows
              const double* abase = reinterpret_cast<const double*>(a.A + (long long)r0 * a.lda + kbase);
              const long long rstride = (C ? 2 : 1) * a.lda;
              double2 av[8][U];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const double* arow = abase + (long long)min(i, nr - 1) * rstride;
                av[i][0] = *reinterpret_cast<const double2*>(arow + off0);
                av[i][1] = *reinterpret_cast<const double2*>(arow + off1);
              }
              const bool interior = (span0 + WC <= r0) && nr == 8;  // every column of the span is left of every row
              T yacc[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = r0 + i;
                const T xr = i < nr ? xs[r - kbase] : H::zero();
                T y = H::zero();
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  const int k = kbase + kw + 32 * VEC * u;
                  if (C) {
                    const bool keep = interior || (i < nr && k < r);
                    const T e0 = keep ? H::make(av[i][u].x, av[i][u].y) : H::zero();
                    y = H::fma(e0, H::make(xv[u].x, xv[u].y), y);
                    zacc[u] = H::cfma(e0, xr, zacc[u]);
                  } else {
                    const double e0 = (interior || (i < nr && k < r)) ? av[i][u].x : 0.0;
                    const double e1 = (interior || (i < nr && k + 1 < r)) ? av[i][u].y : 0.0;
                    y = H::make(::fma(e0, xv[u].x, ::fma(e1, xv[u].y, H::re(y))), 0.0);
                    zacc[2 * u] = H::make(::fma(e0, H::re(xr), H::re(zacc[2 * u])), 0.0);
                    zacc[2 * u + 1] = H::make(::fma(e1, H::re(xr), H::re(zacc[2 * u + 1])), 0.0);
                  }
                }
                yacc[i] = y;
              }
              // Transposing reduction of the 8 row sums over the 32 lanes (9 exchanges instead of 40): after the three
              // halving steps lane L holds row 4 b4 + 2 b3 + b2 (bits of L), two butterfly steps finish the sum.
              {
                const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
                T t4[4], t2[2], t1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const T send = h16 ? yacc[i] : yacc[i + 4], keep = h16 ? yacc[i + 4] : yacc[i];
                  t4[i] = H::add(keep, H::shfl_xor(send, 16));
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const T send = h8 ? t4[i] : t4[i + 2], keep = h8 ? t4[i + 2] : t4[i];
                  t2[i] = H::add(keep, H::shfl_xor(send, 8));
                }
                {
                  const T send = h4 ? t2[0] : t2[1], keep = h4 ? t2[1] : t2[0];
                  t1 = H::add(keep, H::shfl_xor(send, 4));
                }
                t1 = H::add(t1, H::shfl_xor(t1, 2));
                t1 = H::add(t1, H::shfl_xor(t1, 1));
                const int row_i = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                if ((lane & 3) == 0 && row_i < nr) {
                  T* dst = &sm.ysm[warp * 128 + li0 + rb + row_i];
                  *dst = H::add(*dst, t1);
                }
              }
            }
          }
        }
        // this CTA's partial of y[k] for the warp's columns (zeros where it has no rows below k)
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int v2 = 0; v2 < VEC; ++v2) {
            const int k = kbase + kw + 32 * VEC * u + v2;
            if (k >= f && k < n) a.zpart[(long long)b * a.npad + k] = zacc[u * VEC + v2];
          }
      }
    }
    __syncthreads();
    if (prof) tc2 = clock64();
    // per-row: combine the segments (fixed order), publish y_raw, and the row's terms of xx / xy / xa
    if (tid < 128) {
      T cxx = H::zero(), cxy = H::zero(), cxa = H::zero();
      if (tid < rows) {
        const int row = row_of(tid);
        const T x = xs[row - kbase];
        if (!SYM) {
          T y = sm.psum[tid * S];
          for (int s = 1; s < S; ++s) y = H::add(y, sm.psum[tid * S + s]);
          a.ybuf[row] = y;
          cxy = H::cmul(x, y);
        } else {
          // ybuf holds the strictly-lower part plus the diagonal term; the column partials are added by the consumer.
          // x^H A x = 2 Re sum_r conj(x_r) ylow_r + sum_r a_rr |x_r|^2
          T ylow = sm.ysm[tid];
          for (int w = 1; w < HETRD_WARPS; ++w) ylow = H::add(ylow, sm.ysm[w * 128 + tid]);
          const double diag = H::re(a.A[(long long)row * a.lda + row]);
          a.ybuf[row] = H::add(ylow, H::scale(diag, x));
          cxy = H::make(2.0 * H::re(H::cmul(x, ylow)) + diag * H::abs2(x), 0.0);
        }
        if (row != f) cxx = H::make(H::abs2(x), 0.0);
        cxa = H::cmul(x, a.A[(long long)row * a.lda + f]);
      }
      sm.rowred[0][tid] = cxx;
      sm.rowred[1][tid] = cxy;
      sm.rowred[2][tid] = cxa;
    }
    // V^H x, W^H x over the slab: lane jj owns panel column jj
    {
      T av = H::zero(), aw = H::zero();
      if (lane < j)
        for (int t = warp; t < rows; t += HETRD_WARPS) {
          const int row = row_of(t);
          const T x = xs[row - kbase];
          av = H::cfma(H::ldcg(a.V + (long long)row * HNB + lane), x, av);
          aw = H::cfma(H::ldcg(a.W + (long long)row * HNB + lane), x, aw);
        }
      sm.wred[warp][lane] = av;
      sm.wred[warp][HNB + lane] = aw;
    }
    __syncthreads();
    if (warp < 3) {  // fixed-order reduction of the (zero-padded) 128 row slots: 4 per lane, then a butterfly
      T s4 = H::add(H::add(sm.rowred[warp][lane], sm.rowred[warp][lane + 32]),
                    H::add(sm.rowred[warp][lane + 64], sm.rowred[warp][lane + 96]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s4 = H::add(s4, H::shfl_xor(s4, off));
      if (lane == 0) a.part[(long long)warp * a.gpad + b] = s4;
    }
    if (tid >= 96 && tid < 96 + 2 * HNB) {
      const int q = tid - 96;
      T s = sm.wred[0][q];
      for (int w = 1; w < HETRD_WARPS; ++w) s = H::add(s, sm.wred[w][q]);
      a.part[(long long)(3 + q) * a.gpad + b] = s;
    }
    if (prof) tc3 = clock64();
    hetrd_grid_barrier(a.bar, bar_target);  // B2: partial sums, y_raw and x visible everywhere
    if (prof) tc4 = clock64();

    if (SYM) {
      // ---- y[k] += sum over the CTAs of their column partials: every CTA completes a contiguous range of k (coalesced
      // reads of G partial vectors), then one more grid barrier makes the finished y visible --------------------------
      // Rz <= 128 elements per CTA; four threads share an element (a quarter of the CTAs each, 4 loads in flight),
      // combined through shared memory in a fixed order.
      const int Rz = (n1 + G - 1) / G;
      {
        const int t = tid & 127, part = tid >> 7;  // 4 parts x 128 elements
        const long long k = (long long)f + (long long)b * Rz + t;
        T z0 = H::zero(), z1 = H::zero(), z2 = H::zero(), z3 = H::zero();
        if (t < Rz && k < n) {
          const int g_lo = part * ((G + 3) / 4), g_hi = min(G, g_lo + (G + 3) / 4);
          int g = g_lo;
          for (; g + 4 <= g_hi; g += 4) {
            z0 = H::add(z0, H::ldcg(a.zpart + (long long)g * a.npad + k));
            z1 = H::add(z1, H::ldcg(a.zpart + (long long)(g + 1) * a.npad + k));
            z2 = H::add(z2, H::ldcg(a.zpart + (long long)(g + 2) * a.npad + k));
            z3 = H::add(z3, H::ldcg(a.zpart + (long long)(g + 3) * a.npad + k));
          }
          for (; g < g_hi; ++g) z0 = H::add(z0, H::ldcg(a.zpart + (long long)g * a.npad + k));
        }
        sm.psum[tid] = H::add(H::add(z0, z1), H::add(z2, z3));
        __syncthreads();
        if (part == 0 && t < Rz && k < n) {
          const T z = H::add(H::add(sm.psum[t], sm.psum[128 + t]), H::add(sm.psum[256 + t], sm.psum[384 + t]));
          a.ybuf[k] = H::add(H::ldcg(a.ybuf + k), z);
        }
      }
      hetrd_grid_barrier(a.bar, bar_target);  // B3 (symmetric path only)
    }
    // ---- grid reduction of the 3 + 2j sums, identically in every CTA ------------------------------------
    // Warp 0 first issues the loads its scalar phase needs, so that their latency hides behind the reduction.
    T pre_alpha = H::zero(), pre_yraw = H::zero(), pre_aff = H::zero(), pre_rv = H::zero(), pre_rw = H::zero();
    if (warp == 0) {
      pre_alpha = H::ldcg(a.xbuf + f);
      pre_yraw = H::ldcg(a.ybuf + f);
      pre_aff = a.A[(long long)f * a.lda + f];
      if (lane < j) {  // row f of the panel matrices, columns < j (written at least one barrier ago)
        pre_rv = H::ldcg(a.V + (long long)f * HNB + lane);
        pre_rw = H::ldcg(a.W + (long long)f * HNB + lane);
      }
    }
    {
      // Each warp owns the sums q = warp, warp + 16, ...; all of its loads are issued before the first use.
      constexpr int QMAX = (HETRD_NSUM + HETRD_WARPS - 1) / HETRD_WARPS, GMAX = 5;  // G <= 160
      T val[QMAX][GMAX];
#pragma unroll
      for (int t = 0; t < QMAX; ++t) {
        const int q = warp + t * HETRD_WARPS;
        const bool used = q < 3 || (q < HETRD_NSUM && (q < 3 + HNB ? q - 3 < j : q - 3 - HNB < j));
#pragma unroll
        for (int u = 0; u < GMAX; ++u) {
          const int g = lane + 32 * u;
          val[t][u] = (used && g < G) ? H::ldcg(a.part + (long long)q * a.gpad + g) : H::zero();
        }
      }
#pragma unroll
      for (int t = 0; t < QMAX; ++t) {
        const int q = warp + t * HETRD_WARPS;
        T sacc = H::add(H::add(H::add(val[t][0], val[t][1]), H::add(val[t][2], val[t][3])), val[t][4]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sacc = H::add(sacc, H::shfl_xor(sacc, off));
        if (lane == 0 && q < HETRD_NSUM) sm.sums[q] = sacc;
      }
    }
    if (warp == 0 && lane < j) {
      sm.rowV[lane] = pre_rv;
      sm.rowW[lane] = pre_rw;
    }
    __syncthreads();
    // ---- Householder scalars and the derived small vectors (warp 0, redundantly in every CTA) -----------
    if (warp == 0) {
      const T alpha = pre_alpha;
      const double xnorm2 = H::re(sm.sums[0]);
      const bool trivial = (xnorm2 <= 0.0) && (H::im(alpha) == 0.0);
      double beta, tau_re, tau_im, sig_re, sig_im;
      if (trivial) {
        beta = H::re(alpha);
        tau_re = tau_im = sig_re = sig_im = 0.0;
      } else {
        const double nrm = sqrt(H::abs2(alpha) + fmax(xnorm2, 0.0));
        beta = H::re(alpha) >= 0.0 ? -nrm : nrm;
        const double inv_beta = 1.0 / beta;
        tau_re = (beta - H::re(alpha)) * inv_beta;
        tau_im = -H::im(alpha) * inv_beta;
        const double dr = H::re(alpha) - beta, di = H::im(alpha);  // sigma = 1 / (alpha - beta)
        const double inv_den = 1.0 / (dr * dr + di * di);
        sig_re = dr * inv_den;
        sig_im = -di * inv_den;
      }
      const T sigma = H::make(sig_re, sig_im), tau = H::make(tau_re, tau_im);
      // V^H v = sigma (V^H x - beta conj(V[f, :])),  likewise W
      T vv = H::zero(), wv = H::zero();
      if (lane < j) {
        vv = H::mul(sigma, H::sub(sm.sums[3 + lane], H::scale(beta, H::conj(sm.rowV[lane]))));
        wv = H::mul(sigma, H::sub(sm.sums[3 + HNB + lane], H::scale(beta, H::conj(sm.rowW[lane]))));
      }
      sm.Vv[lane] = vv;
      sm.Wv[lane] = wv;
      // y^H v = v^H A v - 2 Re((W^H v)^H (V^H v));  y_f = sigma (y_raw_f - beta a_ff) - V[f,:] (W^H v) - W[f,:] (V^H v)
      double cross = H::re(H::cmul(wv, vv));
      T y0c = H::zero();
      if (lane < j) y0c = H::add(H::mul(sm.rowV[lane], wv), H::mul(sm.rowW[lane], vv));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        cross += __shfl_xor_sync(0xffffffffu, cross, off);
        y0c = H::add(y0c, H::shfl_xor(y0c, off));
      }
      const double aff = H::re(pre_aff);
      const double sig2 = sig_re * sig_re + sig_im * sig_im;
      const double vAv = sig2 * (H::re(sm.sums[1]) - 2.0 * beta * H::re(sm.sums[2]) + beta * beta * aff);
      const double yHv = vAv - 2.0 * cross;
      // w = tau y + alpha2 v,  alpha2 = -1/2 tau conj(tau) (y^H v)   (real: -1/2 tau^2 y^T v)
      const double t2 = tau_re * tau_re + tau_im * tau_im;
      const T alpha2 = H::scale(-0.5 * t2 * yHv, H::make(1.0, 0.0));
      const T yraw_f = pre_yraw;
      const T y0 = H::sub(H::mul(sigma, H::sub(yraw_f, H::make(beta * aff, 0.0))), y0c);
      const T w0 = H::add(H::mul(tau, y0), alpha2);
      if (lane == 0) {
        sm.beta = beta;
        sm.sigma_re = sig_re; sm.sigma_im = sig_im;
        sm.tau_re = tau_re; sm.tau_im = tau_im;
        sm.alpha2_re = H::re(alpha2); sm.alpha2_im = H::im(alpha2);
        sm.rowV[j] = H::make(1.0, 0.0);
        sm.rowW[j] = w0;
        if (b == 0) {
          a.e[c] = beta;
          a.tau[c] = tau;
        }
      }
    }
    __syncthreads();
    if (prof) tc5 = clock64();
    // ---- phase C (+ phase A of the next column): one warp per row, lane q = panel column q ---------------
    {
      const double beta = sm.beta;
      const T sigma = H::make(sm.sigma_re, sm.sigma_im), tau = H::make(sm.tau_re, sm.tau_im);
      const T alpha2 = H::make(sm.alpha2_re, sm.alpha2_im);
      const bool next = j + 1 < a.jb;
      const long long plane = a.npad * 2 * HNB, vplane = a.npad * a.ldv;
      const T wv_l = lane < j ? sm.Wv[lane] : H::zero(), vv_l = lane < j ? sm.Vv[lane] : H::zero();
      const T rw_l = lane < j ? H::conj(sm.rowW[lane]) : H::zero(), rv_l = lane < j ? H::conj(sm.rowV[lane]) : H::zero();
      const T rw_j = H::conj(sm.rowW[j]), rv_j = H::conj(sm.rowV[j]);
      const int gw = b * HETRD_WARPS + warp, nw = G * HETRD_WARPS;
      for (long long r = f + gw; r < n; r += nw) {
        T* vrow = a.V + r * HNB;
        T* wrow = a.W + r * HNB;
        // one coalesced load of the row's panel entries; columns >= j hold stale data and are masked
        const T vq = lane < j ? H::ldcg(vrow + lane) : H::zero();
        const T wq = lane < j ? H::ldcg(wrow + lane) : H::zero();
        const T arf = a.A[r * a.lda + f];          // uniform addresses: one transaction each, issued with the row loads
        const T xr = H::ldcg(a.xbuf + r);
        const T yr = H::ldcg(a.ybuf + r);
        T ys = H::add(H::mul(vq, wv_l), H::mul(wq, vv_l));   // V[r, :j] (W^H v) + W[r, :j] (V^H v)
        T xsum = H::add(H::mul(vq, rw_l), H::mul(wq, rv_l)); // V[r, :j] conj(W[f, :j]) + W[r, :j] conj(V[f, :j])
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          ys = H::add(ys, H::shfl_xor(ys, off));
          xsum = H::add(xsum, H::shfl_xor(xsum, off));
        }
        if (lane == 0) {
          const T v = (r == f) ? H::make(1.0, 0.0) : H::mul(sigma, xr);
          const T y = H::sub(H::mul(sigma, H::sub(yr, H::scale(beta, arf))), ys);
          const T w = H::add(H::mul(tau, y), H::mul(alpha2, v));
          const T xn = H::sub(H::sub(arf, xsum), H::add(H::mul(v, rw_j), H::mul(w, rv_j)));
          vrow[j] = v;
          wrow[j] = w;
          a.PW[r * 2 * HNB + j] = H::re(v);
          a.PW[r * 2 * HNB + HNB + j] = H::re(w);
          a.RW[r * 2 * HNB + j] = H::re(w);
          a.RW[r * 2 * HNB + HNB + j] = H::re(v);
          a.VT[(long long)c * a.ldv + r] = H::re(v);
          a.VR[r * a.ldv + c] = H::re(v);
          if (C) {
            a.PW[plane + r * 2 * HNB + j] = H::im(v);
            a.PW[plane + r * 2 * HNB + HNB + j] = H::im(w);
            a.RW[plane + r * 2 * HNB + j] = H::im(w);
            a.RW[plane + r * 2 * HNB + HNB + j] = H::im(v);
            a.VT[vplane + (long long)c * a.ldv + r] = H::im(v);
            a.VR[vplane + r * a.ldv + c] = -H::im(v);
          }
          if (next) {
            if (r == f) a.d[f] = H::re(xn);
            else a.xbuf[r] = xn;
          }
        }
      }
      // compact-WY factor (CTA 0, after its rows, off the other CTAs' critical path):
      // T[:j, j] = -tau T[:j, :j] (V^H v),  T[j, j] = tau
      if (b == 0 && warp == HETRD_WARPS - 1) {
        T sT = H::zero();
        if (lane < j)
          for (int k = lane; k < j; ++k) sT = H::fma(sm.Tf[lane][k], sm.Vv[k], sT);
        __syncwarp();
        if (lane < j) sm.Tf[lane][j] = H::mul(H::make(-H::re(tau), -H::im(tau)), sT);
        if (lane == 0) sm.Tf[j][j] = tau;
      }
    }
    if (prof) {
      const long long tc6 = clock64();
      a.dbg[0] += tc1 - tc0;  // wait at B1
      a.dbg[1] += tc2 - tc1;  // stage x + A x
      a.dbg[2] += tc3 - tc2;  // row terms, V^H x / W^H x, CTA reduction, publish
      a.dbg[3] += tc4 - tc3;  // wait at B2
      a.dbg[4] += tc5 - tc4;  // grid reduction + Householder scalars
      a.dbg[5] += tc6 - tc5;  // phase C / next x
    }
  }
  // panel done: publish the compact-WY factor
  if (b == 0) {
    __syncthreads();
    T* Tg = a.Tfac + (long long)(a.k0 / HNB) * HNB * HNB;
    for (int t = tid; t < HNB * HNB; t += HETRD_THREADS) Tg[t] = sm.Tf[t / HNB][t % HNB];
  }
}

template <bool C, bool SYM>
size_t hetrd_smem_bytes(int n) {
  return hetrd_xs_bytes<C>(n) + sizeof(HetrdSmem<C, SYM>) + 16;
}

// Rows of the panel operands that the next her2k / back-transformation must see as zero.
__global__ void hetrd_clear_panel_kernel(double* __restrict__ PW, double* __restrict__ RW, long long npad, int planes) {
  const long long total = (long long)planes * npad * 2 * HNB;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    PW[t] = 0.0;
    RW[t] = 0.0;
  }
}

// A[i, j] = scale * A_in[i, j] copied into the work matrix (complex: interleaved in, interleaved out).
template <bool C>
__global__ void hetrd_init_kernel(const double* __restrict__ Ain, long long ld_in, int n, double scale,
                                  typename HS<C>::T* __restrict__ A, long long lda) {
  const long long total = (long long)n * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / n, j = t % n;
    if (C) {
      const double2 v = reinterpret_cast<const double2*>(Ain)[i * ld_in + j];
      reinterpret_cast<double2*>(A)[i * lda + j] = make_double2(scale * v.x, scale * v.y);
    } else {
      reinterpret_cast<double*>(A)[i * lda + j] = scale * Ain[i * ld_in + j];
    }
  }
}

template <bool C>
__global__ void hetrd_last_diag_kernel(const typename HS<C>::T* __restrict__ A, long long lda, int n, double* d) {
  d[n - 1] = HS<C>::re(A[(long long)(n - 1) * lda + n - 1]);
}

// =============================================================================================
// GEMM epilogues (plugged into gemm_kernel, see ops.cuh for the conventions)
// =============================================================================================
// Trailing update after a panel: A[r0 + i, r0 + j] -= sum_k P[i, k] conj(R[j, k]),  P = [V | W], R = [W | V].
template <bool C>
struct OpHer2k {
  struct Params {
    Operand A, B;
    int nt;            // trailing size
    double* Aout;      // &A[r0, r0]; complex: interleaved
    long long lda;     // in elements
    int lower_only;    // update the lower triangle (incl. the diagonal) only: all the symmetric A x path reads
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext;
    t.valid = !p.lower_only || t.n0 < t.m0 + BM;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + warp_m * 32 + 8 * i + (lane >> 2);
      if (row >= p.nt) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = t.n0 + warp_n * 32 + 8 * j + 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (col + e >= p.nt || (p.lower_only && col + e > row)) continue;
          if (C) {
            double2* o = reinterpret_cast<double2*>(p.Aout) + (long long)row * p.lda + col + e;
            double2 v = *o;
            v.x -= acc.r[i][j][e];
            v.y -= acc.i[i][j][e];
            *o = v;
          } else {
            p.Aout[(long long)row * p.lda + col + e] -= acc.r[i][j][e];
          }
        }
      }
    }
  }
};

// Back-transformation, first product (split over K): ws[split][plane][row][0:64) = sum_{k in split} Y[row, k] conj(V[k, j]).
template <bool C>
struct OpBT1 {
  struct Params {
    Operand A, B;      // A = Y (planar), B = VT rows of the panel
    int n_rows;        // eigenvectors
    int k_lo, k_hi;    // K range of the whole product (k_lo multiple of BK)
    int k_per_split;   // multiple of BK
    double* ws;        // [splits][planes][rows_pad][BN]
    long long rows_pad;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = 0;
    t.m0 = blockIdx.x * BM;
    t.k_begin = p.k_lo + blockIdx.y * p.k_per_split;
    t.k_end = min(p.k_hi, t.k_begin + p.k_per_split);
    t.valid = true;  // empty splits still write their zeros
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
    double* wr = p.ws + (long long)blockIdx.y * (C ? 2 : 1) * p.rows_pad * BN;
    double* wi = wr + p.rows_pad * BN;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + warp_m * 32 + 8 * i + (lane >> 2);
      if (row >= p.n_rows) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = warp_n * 32 + 8 * j + 2 * (lane & 3);
        *reinterpret_cast<double2*>(wr + (long long)row * BN + col) = make_double2(acc.r[i][j][0], acc.r[i][j][1]);
        if (C) *reinterpret_cast<double2*>(wi + (long long)row * BN + col) = make_double2(acc.i[i][j][0], acc.i[i][j][1]);
      }
    }
  }
};

// ---- two panels of 32 reflectors merged into one block reflector of 64 (back-transformation only) ----------------
// (I - V1 T1 V1^H)(I - V2 T2 V2^H) = I - [V1 V2] T64 [V1 V2]^H,   T64 = [[T1, -T1 (V1^H V2) T2], [0, T2]].
constexpr int BTW = 2 * HNB;  // width of a merged block reflector

// Cross Gram X_p = V1^H V2 of every pair p (one CTA each): rows [64p, 64p+32) of VT against rows [64p+32, 64p+64),
// contracted over the coordinates.  MODE_COMPLEX yields V1 conj(V2) summed over k, i.e. conj(X); stored as such.
template <bool C>
struct OpCrossGram {
  struct Params {
    Operand A, B;     // both = VT (planar)
    int n;            // coordinates (K extent)
    double* X;        // [pairs][planes][HNB][HNB]
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.m0 = blockIdx.x * BTW;
    t.n0 = blockIdx.x * BTW + HNB;
    t.k_begin = (blockIdx.x * BTW) & ~(BK - 1);
    t.k_end = p.n;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile&, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
    if (warp_m != 0 || warp_n != 0) return;  // the 32 x 32 corner of the tile
    double* xr = p.X + (long long)blockIdx.x * (C ? 2 : 1) * HNB * HNB;
    double* xi = xr + HNB * HNB;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int row = 8 * i + (lane >> 2), col = 8 * j + 2 * (lane & 3) + e;
          xr[row * HNB + col] = acc.r[i][j][e];
          if (C) xi[row * HNB + col] = acc.i[i][j][e];
        }
  }
};

// T64 of every pair from the two panel factors and the cross Gram; one CTA of 32 x 32 threads per pair.
// jb2 = reflectors in the second panel of the LAST pair (0: the pair has a single panel).
template <bool C>
__global__ void __launch_bounds__(1024) bt_merge_t_kernel(const typename HS<C>::T* __restrict__ Tfac,
                                                          const double* __restrict__ X, int npanels,
                                                          typename HS<C>::T* __restrict__ T64) {
  using H = HS<C>;
  using T = typename H::T;
  __shared__ T t1[HNB][HNB + 1], xm[HNB][HNB + 1];  // xm: X, then T1 X
  const int p = blockIdx.x, i = threadIdx.y, j = threadIdx.x;
  const bool has2 = 2 * p + 1 < npanels;
  const T* T1 = Tfac + (long long)(2 * p) * HNB * HNB;
  const T* T2 = Tfac + (long long)(2 * p + 1) * HNB * HNB;
  const double* xr = X + (long long)p * (C ? 2 : 1) * HNB * HNB;
  t1[i][j] = (j >= i) ? T1[i * HNB + j] : H::zero();
  const T t2_ij = (has2 && j >= i) ? T2[i * HNB + j] : H::zero();
  // stored value is sum_k V1[k, i] conj(V2[k, j]) = conj(X[i][j])
  xm[i][j] = has2 ? H::conj(H::make(xr[i * HNB + j], C ? xr[HNB * HNB + i * HNB + j] : 0.0)) : H::zero();
  __syncthreads();
  T s = H::zero();
  for (int k = i; k < HNB; ++k) s = H::fma(t1[i][k], xm[k][j], s);  // (T1 X)[i][j]
  __syncthreads();
  xm[i][j] = s;
  __syncthreads();
  T c12 = H::zero();
  if (has2)
    for (int k = 0; k <= j; ++k) c12 = H::fma(xm[i][k], T2[k * HNB + j], c12);  // (T1 X T2)[i][j]
  T* out = T64 + (long long)p * BTW * BTW;
  out[i * BTW + j] = t1[i][j];
  out[i * BTW + HNB + j] = H::make(-H::re(c12), -H::im(c12));
  out[(HNB + i) * BTW + j] = H::zero();
  out[(HNB + i) * BTW + HNB + j] = t2_ij;
}

// C2[row][j] = sum_i (sum_splits C1[row][i]) T[j][i]   (C2 = C1 T^T, T upper triangular, width jb <= 64), written
// planar with row pitch BTW.  Block: 4 rows x 64 columns.
template <bool C>
__global__ void __launch_bounds__(256) bt_apply_t_kernel(const double* __restrict__ ws, int splits, long long rows_pad,
                                                         int n_rows, const typename HS<C>::T* __restrict__ Tg, int jb,
                                                         double* __restrict__ C2) {
  using H = HS<C>;
  using T = typename H::T;
  __shared__ T c1[4][BTW];
  const int ly = threadIdx.x >> 6, lx = threadIdx.x & 63;
  const long long row = (long long)blockIdx.x * 4 + ly;
  T s = H::zero();
  if (row < n_rows && lx < jb) {
    double re = 0.0, im = 0.0;
    for (int sp = 0; sp < splits; ++sp) {
      const double* base = ws + (long long)sp * (C ? 2 : 1) * rows_pad * BN;
      re += base[row * BN + lx];
      if (C) im += base[rows_pad * BN + row * BN + lx];
    }
    s = H::make(re, im);
  }
  c1[ly][lx] = s;
  __syncthreads();
  if (row < n_rows) {
    T o = H::zero();
    if (lx < jb)
      for (int i = lx; i < jb; ++i) o = H::fma(c1[ly][i], Tg[lx * BTW + i], o);  // sum_{i >= j} C1[., i] T[j][i]
    C2[row * BTW + lx] = H::re(o);
    if (C) C2[rows_pad * BTW + row * BTW + lx] = H::im(o);
  }
}

// Back-transformation, second product: Y[row, k] -= sum_j C2[row, j] V[k, j]  (B planes hold conj(V): VR).
template <bool C>
struct OpBT2 {
  struct Params {
    Operand A, B;      // A = C2 (planar, pitch HNB), B = VR columns of the panel
    int n_rows;        // eigenvectors
    int k_tile0;       // first coordinate tile (multiple of BN) touched by the panel
    int n;             // coordinates
    double* Y;         // planar
    long long ldy, yplane;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = p.k_tile0 + blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + warp_m * 32 + 8 * i + (lane >> 2);
      if (row >= p.n_rows) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = t.n0 + warp_n * 32 + 8 * j + 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (col + e >= p.n) continue;
          const long long o = (long long)row * p.ldy + col + e;
          p.Y[o] -= acc.r[i][j][e];
          if (C) p.Y[p.yplane + o] -= acc.i[i][j][e];
        }
      }
    }
  }
};

// Q_out[l][k] (row-major, complex interleaved) = Y[k][l]  (planar Y, row k = eigenvector k).
__global__ void __launch_bounds__(256) bt_export_complex_kernel(const double* __restrict__ Y, long long ldy,
                                                                long long yplane, int n, double* __restrict__ Qout) {
  __shared__ double tr[32][33], ti[32][33];
  const int k0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int kk = ty; kk < 32; kk += 8) {
    const int k = k0 + kk, l = l0 + tx;
    const bool ok = k < n && l < n;
    tr[kk][tx] = ok ? Y[(long long)k * ldy + l] : 0.0;
    ti[kk][tx] = ok ? Y[yplane + (long long)k * ldy + l] : 0.0;
  }
  __syncthreads();
  for (int ll = ty; ll < 32; ll += 8) {
    const int l = l0 + ll, k = k0 + tx;
    if (l < n && k < n) reinterpret_cast<double2*>(Qout)[(long long)l * n + k] = make_double2(tr[tx][ll], ti[tx][ll]);
  }
}

}  // namespace nls
