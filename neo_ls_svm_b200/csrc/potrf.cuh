// potrf.cuh — hand-written Cholesky factorisation M = U^H U (U upper triangular, row-major: the layout of
// scipy.linalg.cho_factor(..., lower=False)) and the two triangular solves of cho_solve with one right-hand side.
// Replaces scipy.linalg.cho_factor / cho_solve at the reference's _neo_ls_svm.py:177-178 (complex m x m, primal) and
// :313-314 (real n x n, dual).
//
// Blocked right-looking algorithm, block size CNB = 64:
//   chol_diag_kernel   U_kk = chol(M_kk) in shared memory (one CTA)
//   chol_panel_kernel  U_k,rest = U_kk^-H M_k,rest: one thread per column, the column's 64 entries live in shared memory;
//                      also writes the block row transposed and planar (PT[j][p]), the K-contiguous operand of the update
//   OpCholUpdate       M_rest,rest -= U_k,rest^H U_k,rest on the FP64 DMMA GEMM core (upper-triangular tiles only, K = 64)
// The strict lower triangle of the output keeps whatever the input held (scipy's cho_factor does the same).
// Solves: block forward substitution with U^H, block back substitution with U (diag solve + update kernel per block).
#pragma once
#include "hetrd.cuh"  // HS<C> scalar helpers

namespace nls {

constexpr int CNB = 64;

// In place on the diagonal block M[k0:k0+nb, k0:k0+nb] (upper part).  info: first non-positive pivot (1-based), else 0.
template <bool C>
__global__ void __launch_bounds__(256) chol_diag_kernel(typename HS<C>::T* __restrict__ M, long long ld, int k0, int nb,
                                                        int* __restrict__ info) {
  using H = HS<C>;
  using T = typename H::T;
  extern __shared__ __align__(16) unsigned char cd_raw[];
  T(*S)[CNB + 1] = reinterpret_cast<T(*)[CNB + 1]>(cd_raw);
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  for (int i = ty; i < nb; i += 4)
    if (tx < nb) S[i][tx] = (tx >= i) ? M[(long long)(k0 + i) * ld + k0 + tx] : H::zero();
  __syncthreads();
  for (int p = 0; p < nb; ++p) {
    const double piv = H::re(S[p][p]);
    if (!(piv > 0.0)) {
      if (threadIdx.x == 0 && atomicCAS(info, 0, k0 + p + 1) == 0) {}
      return;  // uniform: every thread reads the same pivot
    }
    const double inv = 1.0 / sqrt(piv);
    __syncthreads();
    if (ty == 0 && tx >= p && tx < nb) S[p][tx] = (tx == p) ? H::make(sqrt(piv), 0.0) : H::scale(inv, S[p][tx]);
    __syncthreads();
    for (int i = p + 1 + ty; i < nb; i += 4)
      if (tx >= i && tx < nb) S[i][tx] = H::sub(S[i][tx], H::cmul(S[p][i], S[p][tx]));  // -= conj(U[p][i]) U[p][j]
    __syncthreads();
  }
  for (int i = ty; i < nb; i += 4)
    if (tx >= i && tx < nb) M[(long long)(k0 + i) * ld + k0 + tx] = S[i][tx];
}

// Block row: for every column j >= k0 + nb solve U_kk^H x = M[k0:k0+nb, j]; x overwrites the column.  64 columns per
// CTA (one thread each), the tile and U_kk in shared memory.  PT (planar, pitch CNB): PT[plane][j][p] = x_p.
template <bool C>
__global__ void __launch_bounds__(64) chol_panel_kernel(typename HS<C>::T* __restrict__ M, long long ld, int n, int k0,
                                                        int nb, double* __restrict__ PT, long long pt_plane) {
  using H = HS<C>;
  using T = typename H::T;
  extern __shared__ __align__(16) unsigned char cp_raw[];
  T(*Ukk)[CNB + 1] = reinterpret_cast<T(*)[CNB + 1]>(cp_raw);
  T(*X)[CNB] = reinterpret_cast<T(*)[CNB]>(cp_raw + sizeof(T) * CNB * (CNB + 1));
  const int c = threadIdx.x;
  const long long j = (long long)k0 + nb + (long long)blockIdx.x * 64 + c;
  for (int i = 0; i < nb; ++i)
    if (c < nb) Ukk[i][c] = (c >= i) ? M[(long long)(k0 + i) * ld + k0 + c] : H::zero();
  const bool live = j < n;
  for (int p = 0; p < nb; ++p) X[p][c] = live ? M[(long long)(k0 + p) * ld + j] : H::zero();
  __syncthreads();
  for (int p = 0; p < nb; ++p) {
    const T x = H::scale(1.0 / H::re(Ukk[p][p]), X[p][c]);
    X[p][c] = x;
    for (int q = p + 1; q < nb; ++q) X[q][c] = H::sub(X[q][c], H::cmul(Ukk[p][q], x));  // -= conj(U[p][q]) x_p
  }
  if (live) {
    for (int p = 0; p < nb; ++p) {
      const T x = X[p][c];
      M[(long long)(k0 + p) * ld + j] = x;
      PT[j * CNB + p] = H::re(x);
      if (C) PT[pt_plane + j * CNB + p] = H::im(x);
    }
    for (int p = nb; p < CNB; ++p) {  // a short last block: the update contracts over CNB
      PT[j * CNB + p] = 0.0;
      if (C) PT[pt_plane + j * CNB + p] = 0.0;
    }
  }
}

// Trailing update, upper-triangular tiles: M[r0+i][r0+j] -= sum_p conj(U[p][i]) U[p][j].  MODE_COMPLEX accumulates
// sum_p U[p][i] conj(U[p][j]), the conjugate of what is subtracted.
template <bool C>
struct OpCholUpdate {
  struct Params {
    Operand A, B;   // both = PT rows of the trailing columns
    int nt;         // trailing size
    double* Mout;   // &M[r0][r0]; complex: interleaved
    long long ld;   // in elements
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.n0 = blockIdx.x * BN;
    t.m0 = blockIdx.y * BM;
    t.k_begin = 0;
    t.k_end = p.A.kext;
    t.valid = t.n0 + BN > t.m0;  // the tile touches the upper triangle
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
          if (col + e >= p.nt || col + e < row) continue;
          if (C) {
            double2* o = reinterpret_cast<double2*>(p.Mout) + (long long)row * p.ld + col + e;
            double2 v = *o;
            v.x -= acc.r[i][j][e];
            v.y += acc.i[i][j][e];
            *o = v;
          } else {
            p.Mout[(long long)row * p.ld + col + e] -= acc.r[i][j][e];
          }
        }
      }
    }
  }
};

// ---- triangular solves with one right-hand side --------------------------------------------------------------
// Diagonal block: forward (U_kk^H y = b_k) or backward (U_kk x = y_k), one CTA of 64 threads, column-oriented.
// Thread c first pulls the block entries it will need into registers (forward: column c above the diagonal, backward:
// row c right of it; all loads in flight at once), so the 64 dependent steps run on registers and one shared-memory
// broadcast each instead of a global load per step (40 -> ~8 us per block, 34 blocks per solve at m = 1025).
template <bool C>
__global__ void __launch_bounds__(64) chol_solve_diag_kernel(const typename HS<C>::T* __restrict__ U, long long ld, int k0,
                                                             int nb, int backward, typename HS<C>::T* __restrict__ v) {
  using H = HS<C>;
  using T = typename H::T;
  __shared__ T xs[CNB];
  __shared__ T xp;
  const int c = threadIdx.x;
  T u[CNB];
  double dinv = 0.0;
  if (c < nb) {
    xs[c] = v[k0 + c];
    dinv = 1.0 / H::re(U[(long long)(k0 + c) * ld + k0 + c]);
#pragma unroll
    for (int p = 0; p < CNB; ++p) {
      u[p] = H::zero();
      if (p < nb) {
        if (!backward && p < c) u[p] = U[(long long)(k0 + p) * ld + k0 + c];  // U[p][c]
        if (backward && p > c) u[p] = U[(long long)(k0 + c) * ld + k0 + p];   // U[c][p]
      }
    }
  }
  __syncthreads();
  if (!backward) {
#pragma unroll
    for (int p = 0; p < CNB; ++p) {
      if (p < nb) {
        if (c == p) {
          xs[p] = H::scale(dinv, xs[p]);
          xp = xs[p];
        }
        __syncthreads();
        if (c > p && c < nb) xs[c] = H::sub(xs[c], H::cmul(u[p], xp));  // conj(U[p][c]) y_p
        __syncthreads();
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < CNB; ++q) {
      const int p = CNB - 1 - q;
      if (p < nb) {
        if (c == p) {
          xs[p] = H::scale(dinv, xs[p]);
          xp = xs[p];
        }
        __syncthreads();
        if (c < p) xs[c] = H::sub(xs[c], H::mul(u[p], xp));  // U[c][p] x_p
        __syncthreads();
      }
    }
  }
  if (c < nb) v[k0 + c] = xs[c];
}

// Forward: v[j] -= sum_p conj(U[k0+p][j]) v[k0+p] for j >= k0 + nb.  Backward: v[i] -= sum_p U[i][k0+p] v[k0+p], i < k0.
template <bool C>
__global__ void __launch_bounds__(256) chol_solve_update_kernel(const typename HS<C>::T* __restrict__ U, long long ld,
                                                                int n, int k0, int nb, int backward,
                                                                typename HS<C>::T* __restrict__ v) {
  using H = HS<C>;
  using T = typename H::T;
  __shared__ T xs[CNB];
  if (threadIdx.x < nb) xs[threadIdx.x] = v[k0 + threadIdx.x];
  __syncthreads();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (!backward) {
    const long long j = k0 + nb + t;
    if (j >= n) return;
    T acc = v[j];
    for (int p = 0; p < nb; ++p) acc = H::sub(acc, H::cmul(U[(long long)(k0 + p) * ld + j], xs[p]));
    v[j] = acc;
  } else {
    if (t >= k0) return;
    T acc = v[t];
    const T* row = U + t * ld + k0;
    for (int p = 0; p < nb; ++p) acc = H::sub(acc, H::mul(row[p], xs[p]));
    v[t] = acc;
  }
}


// =============================================================================================
// Triangular inverse X = U^-1 (U upper, row-major; what predict_std contracts with: (gamma C + A)^-1 = U^-1 U^-H,
// _neo_ls_svm.py:467-469 / :473-475).  Block back substitution over the block rows from the bottom,
//     X_k,: = U_kk^-1 (E_k,: - U_k,k+1: X_k+1:,:),
// with the product on the FP64 DMMA GEMM core.  X is built TRANSPOSED and conjugated (Xt[j][p] = conj(X[p][j]),
// planar), which is the K-contiguous operand the next block rows' products need; trtri_export turns it around.
// =============================================================================================
// Planar copy of the upper triangle of U (zeros below the diagonal).
template <bool C>
__global__ void trtri_split_kernel(const typename HS<C>::T* __restrict__ U, long long ld, int n, double* __restrict__ Up,
                                   long long ldu, long long uplane) {
  const long long total = (long long)n * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / n, j = t % n;
    const typename HS<C>::T v = (j >= i) ? U[i * ld + j] : HS<C>::zero();
    Up[i * ldu + j] = HS<C>::re(v);
    if (C) Up[uplane + i * ldu + j] = HS<C>::im(v);
  }
}

// R[i][j] = sum_{p >= k_lo} U[k0+i][p] X[p][j] for the columns j >= j0.
template <bool C>
struct OpTrtriProduct {
  struct Params {
    Operand A, B;      // A = planar U rows k0.., B = Xt rows (planar, conjugated)
    int nb, j0, n, k_lo;
    double* R;         // [planes][CNB][ldr]
    long long ldr, rplane;
  };
  static __device__ __forceinline__ Tile tile(const Params& p) {
    Tile t;
    t.m0 = 0;
    t.n0 = p.j0 + blockIdx.x * BN;
    t.k_begin = p.k_lo;
    t.k_end = p.n;
    t.valid = true;
    return t;
  }
  static __device__ __forceinline__ void epilogue(const Params& p, const Tile& t, Acc& acc, int warp_m, int warp_n,
                                                  int lane, uint8_t*) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = warp_m * 32 + 8 * i + (lane >> 2);
      if (row >= p.nb) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = t.n0 + warp_n * 32 + 8 * j + 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (col + e >= p.n) continue;
          p.R[(long long)row * p.ldr + col + e] = acc.r[i][j][e];
          if (C) p.R[p.rplane + (long long)row * p.ldr + col + e] = acc.i[i][j][e];
        }
      }
    }
  }
};

// Block row k of X: for every column j >= k0 solve U_kk x = e_j[k0:k0+nb] - R[:, j] (R = 0 for j < k0 + nb) and store
// Xt[j][k0 + i] = conj(x_i).  64 columns per CTA, one thread each.
template <bool C>
__global__ void __launch_bounds__(64) trtri_block_kernel(const double* __restrict__ Up, long long ldu, long long uplane,
                                                         int n, int k0, int nb, const double* __restrict__ R, long long ldr,
                                                         long long rplane, double* __restrict__ Xt) {
  using H = HS<C>;
  using T = typename H::T;
  extern __shared__ __align__(16) unsigned char tb_raw[];
  T(*Ukk)[CNB + 1] = reinterpret_cast<T(*)[CNB + 1]>(tb_raw);
  T(*X)[CNB] = reinterpret_cast<T(*)[CNB]>(tb_raw + sizeof(T) * CNB * (CNB + 1));
  const int c = threadIdx.x;
  const long long j = (long long)k0 + (long long)blockIdx.x * 64 + c;
  for (int i = 0; i < nb; ++i)
    if (c < nb) {
      const long long o = (long long)(k0 + i) * ldu + k0 + c;
      Ukk[i][c] = H::make(Up[o], C ? Up[uplane + o] : 0.0);
    }
  const bool live = j < n;
  for (int i = 0; i < nb; ++i) {
    T rhs = H::make((j == k0 + i) ? 1.0 : 0.0, 0.0);
    if (live && j >= k0 + nb) rhs = H::sub(rhs, H::make(R[(long long)i * ldr + j], C ? R[rplane + (long long)i * ldr + j] : 0.0));
    X[i][c] = rhs;
  }
  __syncthreads();
  for (int p = nb - 1; p >= 0; --p) {
    // 1 / U[p][p] for a complex diagonal entry (Cholesky factors have a real one, a general triangular U need not)
    const T d = Ukk[p][p];
    const double den = H::abs2(d);
    const T x = H::mul(X[p][c], H::make(H::re(d) / den, -H::im(d) / den));
    X[p][c] = x;
    for (int q = 0; q < p; ++q) X[q][c] = H::sub(X[q][c], H::mul(Ukk[q][p], x));
  }
  if (live)
    for (int i = 0; i < nb; ++i) {
      const long long o = j * ldu + k0 + i;
      Xt[o] = H::re(X[i][c]);
      if (C) Xt[uplane + o] = -H::im(X[i][c]);
    }
}

// Bout[p][j] = X[p][j] = conj(Xt[j][p]); 32 x 32 tiles through shared memory.
template <bool C>
__global__ void __launch_bounds__(256) trtri_export_kernel(const double* __restrict__ Xt, long long ldu, long long uplane,
                                                           int n, typename HS<C>::T* __restrict__ Bout) {
  __shared__ double tr[32][33], ti[32][33];
  const int j0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int jj = ty; jj < 32; jj += 8) {
    const int j = j0 + jj, p = p0 + tx;
    const bool ok = j < n && p < n;
    tr[jj][tx] = ok ? Xt[(long long)j * ldu + p] : 0.0;
    ti[jj][tx] = (ok && C) ? Xt[uplane + (long long)j * ldu + p] : 0.0;
  }
  __syncthreads();
  for (int pp = ty; pp < 32; pp += 8) {
    const int p = p0 + pp, j = j0 + tx;
    if (p < n && j < n) Bout[(long long)p * n + j] = HS<C>::make(tr[tx][pp], -ti[tx][pp]);
  }
}

}  // namespace nls
