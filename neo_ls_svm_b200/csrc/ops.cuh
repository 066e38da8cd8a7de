// ops.cuh — per-stage tile mappings and fused epilogues plugged into gemm_kernel (gemm_core.cuh).
#pragma once
#include "gemm_core.cuh"

namespace nls {

// Coordinates of accumulator element (i, j, e) of a warp tile inside the CTA tile.
__device__ __forceinline__ int acc_row(int warp_m, int i, int lane) { return warp_m * 32 + 8 * i + (lane >> 2); }
__device__ __forceinline__ int acc_col(int warp_n, int j, int lane) { return warp_n * 32 + 8 * j + 2 * (lane & 3); }

// Shared tail of the two LOO sweeps (primal and dual): acc.r holds the unclipped LOO residuals of the
// CTA tile (rows x gammas).  Applies the classifier clip (_neo_ls_svm.py:153-155 / :290-292), takes
// |.| and reduces s-weighted over rows in a fixed order:
//   part[m_tile][0][g] = sum_i s_i |e_ig|,  [1] = sum_i s_i (|e_ig| >= 1),  [2] = sum_i s_i max(0, |e_ig| - 1).
__device__ __forceinline__ void loo_reduce(Acc& acc, const Tile& t, int n_rows, int G, const double* __restrict__ y,
                                           const double* __restrict__ s, int is_classifier, double* __restrict__ part,
                                           int warp_m, int warp_n, int lane, uint8_t* scratch) {
  double e_abs[4][2], e_cnt[4][2], e_hng[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) e_abs[j][e] = e_cnt[j][e] = e_hng[j][e] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = t.m0 + acc_row(warp_m, i, lane);
    if (row >= n_rows) continue;
    const double yi = y[row];
    const double si = s[row];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double loo = acc.r[i][j][e];
        if (is_classifier) {
          if ((yi > 0.0 && loo > 0.0) || (yi < 0.0 && loo < 0.0)) loo = 0.0;
        }
        const double a = fabs(loo);
        e_abs[j][e] += si * a;
        if (is_classifier) {
          e_cnt[j][e] += (a >= 1.0) ? si : 0.0;
          e_hng[j][e] += si * fmax(0.0, a - 1.0);
        }
      }
  }
  // Reduce over the 8 row-lanes of the warp (lanes sharing lane%4), fixed butterfly order.
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        e_abs[j][e] += __shfl_xor_sync(0xffffffffu, e_abs[j][e], off);
        e_cnt[j][e] += __shfl_xor_sync(0xffffffffu, e_cnt[j][e], off);
        e_hng[j][e] += __shfl_xor_sync(0xffffffffu, e_hng[j][e], off);
      }
  double* red = reinterpret_cast<double*>(scratch);  // [3][4 warp_m][64 cols]
  if ((lane >> 2) == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = acc_col(warp_n, j, lane) + e;
        red[(0 * 4 + warp_m) * BN + c] = e_abs[j][e];
        red[(1 * 4 + warp_m) * BN + c] = e_cnt[j][e];
        red[(2 * 4 + warp_m) * BN + c] = e_hng[j][e];
      }
  }
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < 3 * BN) {
    const int q = tid / BN, c = tid % BN;
    const int g = t.n0 + c;
    if (g < G) {
      const double* r = red + (q * 4) * BN + c;
      const double sum = ((r[0] + r[BN]) + r[2 * BN]) + r[3 * BN];
      part[((long long)blockIdx.y * 3 + q) * G + g] = sum;
    }
  }
}

// =============================================================================================
// Stage 1: z = Xc . W^T-tile, epilogue (cos z, sin z)/sqrt(D) in one of three layouts.
//   reference: _affine_feature_map.py:81-89 (z) and _feature_maps.py:201-202 (exp(-1j z)/sqrt(D)).
// =============================================================================================
enum { FM_PLANAR = 0, FM_TRANSPOSED = 1, FM_COMPLEX = 2, FM_NONE = 3 };

// Optional fused row dot products of the feature-map epilogue: for up to two coefficient vectors q,
//     dot_q[row] = sum_l ( c_l x_q[l * stride] + ysign_q s_l y_q[l * stride] )        (c, s = cos, sin / sqrt(D))
// accumulated over the CTA tile's 64 features and written as one partial per feature tile,
//     out[(feature_tile * 2 + q) * ld + row],
// to be summed over the feature tiles in a fixed order by the consumer.  With (x, y) = (Re b, Im b), ysign = +1 this
// is Re(phi b) (_neo_ls_svm.py:664); with (Re q, Im q, +1) and (Im q, Re q, -1) it is Re and Im of phi q.  It replaces
// GEMV kernels that re-read the 16 D bytes per row of the feature chunk from HBM.
struct RowDots {
  const double* x[2];
  const double* y[2];
  double ysign[2];
  int stride;
  int count;      // 0 (off), 1 or 2
  double* out;
  long long ld;
};

struct OpFeatureMap {
  struct Params {
    Operand A, B;      // A = centred rows (n_rows x d), B = W^T (D x d)
    int n_rows, D;
    int layout;
    double inv_sqrt_D;
    const double* row_scale;  // optional per-row factor (sample weight s_i), may be null
    double* out;
    long long ld;             // FM_PLANAR: row pitch (>= 2*Dp); FM_TRANSPOSED: pitch of a feature row
    int plane_stride;         // FM_PLANAR: column offset of the sin plane; FM_TRANSPOSED: row offset
    RowDots dots;             // fused row dot products (count = 0: none); FM_NONE stores nothing else
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t* scratch) {
    const int ndots = p.dots.count;
    double dsum[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) dsum[i][0] = dsum[i][1] = 0.0;
    // Coefficients of the fused dot products for the tile's 64 columns, staged once in shared memory as
    // {x0, ysign0 y0, x1, ysign1 y1} per column (registers would push the kernel past 128 per thread and with that
    // from two resident CTAs per SM to one: measured 94 -> 142 ms per fit for the three feature-map passes).
    double4* coef = reinterpret_cast<double4*>(scratch);  // [BN]
    if (ndots) {
      const int tid = threadIdx.x;
      if (tid < BN) {
        const int col = t.n0 + tid;
        const bool in = col < p.D;
        const long long l = (long long)col * p.dots.stride;
        double4 q;
        q.x = in ? p.dots.x[0][l] : 0.0;
        q.y = in ? p.dots.ysign[0] * p.dots.y[0][l] : 0.0;
        q.z = (in && ndots > 1) ? p.dots.x[1][l] : 0.0;
        q.w = (in && ndots > 1) ? p.dots.ysign[1] * p.dots.y[1][l] : 0.0;
        coef[tid] = q;
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + acc_row(warp_m, i, lane);
      if (row >= p.n_rows) continue;
      const double w = p.row_scale ? p.row_scale[row] * p.inv_sqrt_D : p.inv_sqrt_D;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = t.n0 + acc_col(warp_n, j, lane);
        double c[2], s[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sincos(acc.r[i][j][e], &s[e], &c[e]);
          c[e] *= w;
          s[e] *= w;
        }
        if (ndots) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double4 q = coef[acc_col(warp_n, j, lane) + e];
            dsum[i][0] += c[e] * q.x + s[e] * q.y;
            dsum[i][1] += c[e] * q.z + s[e] * q.w;
          }
        }
        if (p.layout == FM_NONE) {
          // dot products only: the chunk itself is not needed by any later kernel
        } else if (p.layout == FM_PLANAR) {
          double* o = p.out + (long long)row * p.ld + col;
          if (col + 1 < p.D) {
            *reinterpret_cast<double2*>(o) = make_double2(c[0], c[1]);
            *reinterpret_cast<double2*>(o + p.plane_stride) = make_double2(s[0], s[1]);
          } else if (col < p.D) {
            o[0] = c[0];
            o[p.plane_stride] = s[0];
          }
        } else if (p.layout == FM_TRANSPOSED) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (col + e < p.D) {
              p.out[(long long)(col + e) * p.ld + row] = c[e];
              p.out[(long long)(col + e + p.plane_stride) * p.ld + row] = s[e];
            }
        } else {  // FM_COMPLEX: phi = (c, -s) interleaved, row pitch (D+1) complex
          double* o = p.out + ((long long)row * (p.D + 1) + col) * 2;
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (col + e < p.D) {
              o[2 * e] = c[e];
              o[2 * e + 1] = -s[e];
            }
          if (col == 0) {  // constant feature, _feature_maps.py:203
            double* last = p.out + ((long long)row * (p.D + 1) + p.D) * 2;
            last[0] = 1.0;
            last[1] = 0.0;
          }
        }
      }
    }
    if (ndots) {  // uniform over the grid
      // Row sums over the tile's 64 features in a fixed order: the 4 lanes of a row (butterfly), then the two
      // column warps through shared memory.
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          dsum[i][q] += __shfl_xor_sync(0xffffffffu, dsum[i][q], 1);
          dsum[i][q] += __shfl_xor_sync(0xffffffffu, dsum[i][q], 2);
        }
      double* red = reinterpret_cast<double*>(scratch) + 4 * BN;  // [2 warp_n][2 q][BM rows], after the coefficients
      if ((lane & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = acc_row(warp_m, i, lane);
          red[(warp_n * 2 + 0) * BM + r] = dsum[i][0];
          red[(warp_n * 2 + 1) * BM + r] = dsum[i][1];
        }
      }
      __syncthreads();
      const int tid = threadIdx.x;
      if (tid < ndots * BM) {
        const int q = tid / BM, r = tid % BM, row = t.m0 + r;
        if (row < p.n_rows)
          p.dots.out[((long long)blockIdx.x * 2 + q) * p.dots.ld + row] = red[(0 * 2 + q) * BM + r] + red[(1 * 2 + q) * BM + r];
      }
    }
  }
};

// =============================================================================================
// Stage 2: Hermitian Gram of the weighted, transposed feature chunk Psi^T (2*Dp x rows):
//   R = C^T C + S^T S = Re A,   I = S^T C - C^T S = Im A    (upper-triangular tiles only),
// split over the chunk's rows; each (split, tile) CTA owns its slot of the partial-sum workspace and
// accumulates into it chunk after chunk (fixed order => bitwise reproducible).
//   reference: _neo_ls_svm.py:112-114.
// =============================================================================================
struct OpGram {
  struct Params {
    Operand A, B;   // both = Psi^T; A box 128 rows, B box 64 rows
    int D;
    int n_tiles_m;  // ceil(D/128)
    int n_tiles_n;  // ceil(D/64)
    int rows;       // valid rows (K extent) in this chunk
    int k_per_split;
    double* ws;     // [split][2][D][D]
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    int idx = blockIdx.x, kb = 0;
    // Row block kb owns column blocks lb >= 2*kb (tiles touching the upper triangle).
    while (kb < p.n_tiles_m && idx >= p.n_tiles_n - 2 * kb) {
      idx -= p.n_tiles_n - 2 * kb;
      ++kb;
    }
    t.valid = kb < p.n_tiles_m;
    t.m0 = kb * BM;
    t.n0 = (2 * kb + idx) * BN;
    t.k_begin = blockIdx.y * p.k_per_split;
    t.k_end = min(p.rows, t.k_begin + p.k_per_split);
    if (t.k_begin >= t.k_end) t.valid = false;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
    double* wr = p.ws + (long long)blockIdx.y * 2 * p.D * p.D;
    double* wi = wr + (long long)p.D * p.D;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + acc_row(warp_m, i, lane);
      if (row >= p.D) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = t.n0 + acc_col(warp_n, j, lane);
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (col + e < p.D && col + e >= row) {
            const long long o = (long long)row * p.D + col + e;
            wr[o] += acc.r[i][j][e];
            wi[o] += acc.i[i][j][e];
          }
      }
    }
  }
};

// =============================================================================================
// Stage 4a: T = phi Q for a chunk (complex GEMM on planar [C|S]), epilogue
//   P = Re(T v),  U = |T|^2 inv_c      reference: _neo_ls_svm.py:134, :137 (single-T form)
// With xa = C, xb = S, ya = Re Q^T, yb = Im Q^T:  R = Re T,  I = -Im T  (before the constant-feature
// bias Q[D, k], which is added here).
// =============================================================================================
struct OpProject {
  struct Params {
    Operand A, B;         // A = Psi chunk (rows x 2Dp), B = Q^T planes (2Np x Dp)
    int n_rows, m;
    const double* bias_r;  // Re Q[D, k]
    const double* bias_i;  // Im Q[D, k]
    const double* v_r;
    const double* v_i;
    double inv_c;
    double* P;
    double* U;
    long long ld;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext < p.B.kext ? p.A.kext : p.B.kext;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = t.n0 + acc_col(warp_n, j, lane);
      double br[2], bi[2], vr[2], vi[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = col + e < p.m;
        br[e] = ok ? p.bias_r[col + e] : 0.0;
        bi[e] = ok ? p.bias_i[col + e] : 0.0;
        vr[e] = ok ? p.v_r[col + e] : 0.0;
        vi[e] = ok ? p.v_i[col + e] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = t.m0 + acc_row(warp_m, i, lane);
        if (row >= p.n_rows) continue;
        double pv[2], uv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double tr = acc.r[i][j][e] + br[e];
          const double ti = bi[e] - acc.i[i][j][e];
          pv[e] = tr * vr[e] - ti * vi[e];
          uv[e] = (tr * tr + ti * ti) * p.inv_c;
        }
        const long long o = (long long)row * p.ld + col;
        if (col + 1 < p.m) {
          *reinterpret_cast<double2*>(p.P + o) = make_double2(pv[0], pv[1]);
          *reinterpret_cast<double2*>(p.U + o) = make_double2(uv[0], uv[1]);
        } else if (col < p.m) {
          p.P[o] = pv[0];
          p.U[o] = uv[0];
        }
      }
    }
  }
};

// =============================================================================================
// Stage 4b: num = P r, den = s^2 (U r) with r[k, g] = 1/(gamma_g + lam_k); fused LOO residual,
// classifier clip, |.|, and the s-weighted reduction over rows.  The n x G matrices never exist.
//   reference: _neo_ls_svm.py:147-161.
// Output: part[m_tile][3][G] (fixed-order partial sums, reduced by sweep_reduce_kernel).
// =============================================================================================
struct OpSweep {
  struct Params {
    Operand A, B;   // A = [P ; U] stacked (plane_drow = chunk capacity), B = r^T (G x ldp)
    int n_rows, G;
    const double* y;
    const double* s;
    int is_classifier;
    double* part;   // [gridDim.y][3][G]
    double* den_out;  // optional stash of U r (= sigma2_i at every gamma), row pitch den_ld; may be null
    long long den_ld;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t* scratch) {
    if (p.den_out) {
      // Stash sigma2_i(gamma_g) = (U r)_ig so that the per-row outputs at the selected gamma need no
      // second projection pass (the n x G matrix costs HBM capacity, not time: 8 B per MMA'd 2m flops).
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = t.m0 + acc_row(warp_m, i, lane);
        if (row >= p.n_rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = t.n0 + acc_col(warp_n, j, lane);
          double* o = p.den_out + (long long)row * p.den_ld + col;
          if (col + 1 < p.G && (p.den_ld & 1) == 0) {
            *reinterpret_cast<double2*>(o) = make_double2(acc.i[i][j][0], acc.i[i][j][1]);
          } else {
            if (col < p.G) o[0] = acc.i[i][j][0];
            if (col + 1 < p.G) o[1] = acc.i[i][j][1];
          }
        }
      }
    }
    // Leave-one-out residuals in place of the numerators, then the shared weighted reduction.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + acc_row(warp_m, i, lane);
      const double yi = row < p.n_rows ? p.y[row] : 0.0;
      const double si = row < p.n_rows ? p.s[row] : 0.0;
      const double s2 = si * si;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) acc.r[i][j][e] = (acc.r[i][j][e] - yi) / (1.0 - s2 * acc.i[i][j][e]);
    }
    loo_reduce(acc, t, p.n_rows, p.G, p.y, p.s, p.is_classifier, p.part, warp_m, warp_n, lane, scratch);
  }
};

// =============================================================================================
// Stage 4c / 5b: sigma2_i = sum_k |(phi B)_ik|^2 w_k — same mainloop as OpProject, epilogue reduces
// over the tile's columns.   reference: _neo_ls_svm.py:184 and :467-469 (eigenbasis / U^-1 form).
// Output: part[n_tile][row] (reduced over n_tile in fixed order by rowsum_reduce_kernel).
// =============================================================================================
template <bool COMPLEX>
struct OpVarianceT {
  struct Params {
    Operand A, B;
    int n_rows, m;
    const double* bias_r;
    const double* bias_i;
    const double* w;
    double* part;     // [gridDim.x][part_ld]
    long long part_ld;
    int b_upper;      // B is upper triangular (B = U^-1): column k only needs rows l <= k of the contraction
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext < p.B.kext ? p.A.kext : p.B.kext;
    if (p.b_upper && t.n0 + BN < t.k_end) t.k_end = t.n0 + BN;  // halves the flops of predict_std
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t* scratch) {
    double rs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = t.n0 + acc_col(warp_n, j, lane);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = col + e < p.m;
        const double br = (ok && COMPLEX) ? p.bias_r[col + e] : 0.0;
        const double bi = (ok && COMPLEX) ? p.bias_i[col + e] : 0.0;
        const double w = ok ? p.w[col + e] : 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double tr = acc.r[i][j][e] + br;
          const double ti = COMPLEX ? bi - acc.i[i][j][e] : 0.0;
          rs[i] += (tr * tr + ti * ti) * w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
      rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
    }
    double* red = reinterpret_cast<double*>(scratch);  // [2 warp_n][128 rows]
    if ((lane & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) red[warp_n * BM + acc_row(warp_m, i, lane)] = rs[i];
    }
    __syncthreads();
    const int tid = threadIdx.x;
    if (tid < BM) {
      const int row = t.m0 + tid;
      if (row < p.n_rows) p.part[(long long)blockIdx.x * p.part_ld + row] = red[tid] + red[BM + tid];
    }
  }
};

using OpVariance = OpVarianceT<true>;   // primal: |phi B|^2 w
using OpRowQuad = OpVarianceT<false>;  // dual:   (K B^T)^2 w

// =============================================================================================
// Dual path (reference _optimize_alpha_gamma, _neo_ls_svm.py:252-323, einsum-free restatement of
// SURVEY.md §8c).  With SQ = sn*Q, r[k,g] = 1/(gamma_g + lam_k):
//   OpDualCoef  : [alpha_mat ; SQ*SQ] r  ->  alpha_loo (n x G, and transposed), hdiag (0 -> eps, :281)
//   OpDualCross : F SQ                    ->  G1 = F0 SQ (F0 = F with zero diagonal), M = G1 * SQ
//   OpDualSweep : F alpha_loo (plane a) and M r (plane b)  ->  yhat_loo (:286) and the error sums
// =============================================================================================
struct OpDualCoef {
  struct Params {
    Operand A, B;
    int n, G;
    double* aloo;   // n x G
    double* alooT;  // G x ldT
    long long ldT;
    double* hd;     // n x G
    double eps;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
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
      const int row = t.m0 + acc_row(warp_m, i, lane);
      if (row >= p.n) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = t.n0 + acc_col(warp_n, j, lane) + e;
          if (col >= p.G) continue;
          const double a = acc.r[i][j][e];
          const double h = acc.i[i][j][e];
          p.aloo[(long long)row * p.G + col] = a;
          p.alooT[(long long)col * p.ldT + row] = a;
          p.hd[(long long)row * p.G + col] = (h == 0.0) ? p.eps : h;
        }
    }
  }
};

struct OpDualCross {
  struct Params {
    Operand A, B;
    int n;
    const double* SQ;
    long long ld;
    double fdiag;
    double* G1;
    double* M;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
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
      const int row = t.m0 + acc_row(warp_m, i, lane);
      if (row >= p.n) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = t.n0 + acc_col(warp_n, j, lane) + e;
          if (col >= p.n) continue;
          const long long o = (long long)row * p.ld + col;
          const double sq = p.SQ[o];
          const double g1 = acc.r[i][j][e] - p.fdiag * sq;
          p.G1[o] = g1;
          p.M[o] = g1 * sq;
        }
    }
  }
};

struct OpDualSweep {
  struct Params {
    Operand A, B;  // A = [F ; M] stacked, B = [alpha_loo^T ; r^T] stacked
    int n_rows, G;
    const double* y;
    const double* s;
    int is_classifier;
    double* part;
    const double* aloo;
    const double* hd;
    double fdiag;
    double* yhat_loo;  // n x G
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t* scratch) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = t.m0 + acc_row(warp_m, i, lane);
      const bool rok = row < p.n_rows;
      const double yi = rok ? p.y[row] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = t.n0 + acc_col(warp_n, j, lane) + e;
          double yh = 0.0;
          if (rok && col < p.G) {
            const long long o = (long long)row * p.G + col;
            const double a = p.aloo[o];
            yh = (-acc.i[i][j][e] / p.hd[o]) * a + (acc.r[i][j][e] - p.fdiag * a);
            p.yhat_loo[o] = yh;
          }
          acc.r[i][j][e] = yh - yi;
        }
    }
    loo_reduce(acc, t, p.n_rows, p.G, p.y, p.s, p.is_classifier, p.part, warp_m, warp_n, lane, scratch);
  }
};

// =============================================================================================
// Plain tile store (used by the dual path: kernel matrix products) with an optional exp epilogue:
//   out[row, col] = R                         (EXP = false)
//   out[row, col] = exp(R - 0.5 (na_row + nb_col)) + add   (EXP = true; R = xa.xb^T)
//   reference: sklearn rbf_kernel(gamma=0.5) at _neo_ls_svm.py:261, :474, :669.
// =============================================================================================
template <bool EXP>
struct OpStore {
  struct Params {
    Operand A, B;
    int n_rows, n_cols;
    const double* norm_a;
    const double* norm_b;
    double add;
    int zero_diag_to;  // if >= 0 and row == col: store exp(0) + add exactly (distance forced to 0)
    double* out;
    long long ld;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
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
      const int row = t.m0 + acc_row(warp_m, i, lane);
      if (row >= p.n_rows) continue;
      const double na = EXP ? p.norm_a[row] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = t.n0 + acc_col(warp_n, j, lane);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (col + e >= p.n_cols) continue;
          double v = acc.r[i][j][e];
          if (EXP) {
            double d2 = na - 2.0 * v + p.norm_b[col + e];
            if (d2 < 0.0) d2 = 0.0;
            if (p.zero_diag_to >= 0 && row == col + e) d2 = 0.0;
            v = exp(-0.5 * d2) + p.add;
          }
          p.out[(long long)row * p.ld + col + e] = v;
        }
      }
    }
  }
};

}  // namespace nls
