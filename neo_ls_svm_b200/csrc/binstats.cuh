// binstats.cuh — per-bin weighted median and weighted mean absolute deviation of every feature column,
// the O(n d log n) part of the supervised affine pre-pass (reference: _affine_normalizer.py:81-88 via
// _weighted_quantile.py:35-77, which argsorts every column of every target bin on the host).
//
// No sort: the rows are grouped by bin (a permutation built on the host), and the weighted median of all
// (bin, column) pairs is found simultaneously by bisection on the order-preserving 64-bit key of the
// doubles: 64 passes over X, each an HBM-bound masked weighted sum with fixed-order reductions
// (bitwise reproducible).  One more pass collects what the reference's two linear interpolations need
// around the crossing point (neighbouring values, cumulative weights, tie-group size), one more the MAD.
// 66 passes x 2 GB = 135 GB at n = 4M, d = 64: ~25 ms of HBM time instead of ~35 s of host argsort.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nls {

constexpr int BS_COLS = 256;     // columns per CTA (8 per lane)
constexpr int BS_K = BS_COLS / 32;
constexpr int BS_TILE_ROWS = 2048;

__device__ __forceinline__ unsigned long long dkey(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double keyd(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct BsTile {
  int bin, r0, r1, pad;
};

// partial[tile, col] = sum over the tile's rows of w * [key(x) <= mid[bin, col]]
// K = columns per lane (32 K columns per CTA); templated so that narrow matrices do not pay registers for
// 256 columns.  Four rows are in flight per warp (independent gathers through perm).
template <int K>
__global__ void __launch_bounds__(256) bs_count_kernel(const double* __restrict__ X, int d,
                                                       const long long* __restrict__ perm,
                                                       const double* __restrict__ w, const BsTile* __restrict__ tiles,
                                                       const unsigned long long* __restrict__ mid,
                                                       double* __restrict__ partial) {
  __shared__ double red[8][32 * K];
  const BsTile t = tiles[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * 32 * K;
  unsigned long long m[K];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int col = col0 + lane + 32 * k;
    m[k] = col < d ? mid[(long long)t.bin * d + col] : 0ull;
    acc[k] = 0.0;
  }
  for (int r = t.r0 + warp; r < t.r1; r += 32) {
    double xv[4][K], wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = r + 8 * u;
      const bool ok = rr < t.r1;
      const double* x = X + (ok ? perm[rr] : perm[r]) * d;
      wv[u] = ok ? w[rr] : 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int col = col0 + lane + 32 * k;
        xv[u][k] = col < d ? x[col] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += (dkey(xv[u][k]) <= m[k]) ? wv[u] : 0.0;  // accumulated in row order
  }
#pragma unroll
  for (int k = 0; k < K; ++k) red[warp][lane + 32 * k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * K; c += 256) {
    const int col = col0 + c;
    if (col < d) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][c];
      partial[(long long)blockIdx.x * d + col] = s;
    }
  }
}

// One thread per (bin, col): F = sum of the bin's tile partials (fixed order), then one bisection step
// towards the smallest key v with F(v) > threshold[bin] (0.5 * total weight when `thresh` is null).
// iter < 0: initialise (F is the bin's total weight).
__global__ void bs_step_kernel(const double* __restrict__ partial, const int2* __restrict__ bin_tiles, int nbins, int d,
                               int iter, const double* __restrict__ thresh, double* __restrict__ wtot,
                               unsigned long long* __restrict__ lo, unsigned long long* __restrict__ hi,
                               unsigned long long* __restrict__ mid) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nbins * d) return;
  const int b = e / d, col = e % d;
  double F = 0.0;
  for (int t = bin_tiles[b].x; t < bin_tiles[b].y; ++t) F += partial[(long long)t * d + col];
  if (iter < 0) {
    wtot[e] = F;
    lo[e] = 0ull;
    hi[e] = ~0ull;
  } else {
    const double cut = thresh ? thresh[b] : 0.5 * wtot[e];
    if (F > cut) hi[e] = mid[e]; else lo[e] = mid[e] + 1ull;
  }
  mid[e] = lo[e] + ((hi[e] - lo[e]) >> 1);
}

// Statistics around the crossing value v* (key in `mid`), per tile:
//   0: W_lt = sum w [x < v*]   1: W_eq = sum w [x == v*]   2: n_eq   3: pred = max x < v* (as key, 0 if none)
//   4: succ = min x > v* (as key, ~0 if none)   5: first position r with x == v*   6: its weight
__global__ void __launch_bounds__(256) bs_stats_kernel(const double* __restrict__ X, int d,
                                                       const long long* __restrict__ perm,
                                                       const double* __restrict__ w, const BsTile* __restrict__ tiles,
                                                       const unsigned long long* __restrict__ vkey,
                                                       double* __restrict__ partial /* [tile][7][d] */) {
  __shared__ double red[8][BS_COLS];
  __shared__ unsigned long long redk[8][BS_COLS];
  const BsTile t = tiles[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * BS_COLS;
  unsigned long long v[BS_K], pred[BS_K], succ[BS_K], fpos[BS_K];
  double wlt[BS_K], weq[BS_K], neq[BS_K], fw[BS_K];
#pragma unroll
  for (int k = 0; k < BS_K; ++k) {
    const int col = col0 + lane + 32 * k;
    v[k] = col < d ? vkey[(long long)t.bin * d + col] : 0ull;
    pred[k] = 0ull;
    succ[k] = ~0ull;
    fpos[k] = ~0ull;
    wlt[k] = weq[k] = neq[k] = fw[k] = 0.0;
  }
  for (int r = t.r0 + warp; r < t.r1; r += 8) {
    const double* x = X + perm[r] * d;
    const double wr = w[r];
#pragma unroll
    for (int k = 0; k < BS_K; ++k) {
      const int col = col0 + lane + 32 * k;
      if (col >= d) continue;
      const unsigned long long key = dkey(x[col]);
      if (key < v[k]) {
        wlt[k] += wr;
        pred[k] = key > pred[k] ? key : pred[k];
      } else if (key == v[k]) {
        weq[k] += wr;
        neq[k] += 1.0;
        if ((unsigned long long)r < fpos[k]) {
          fpos[k] = (unsigned long long)r;
          fw[k] = wr;
        }
      } else {
        succ[k] = key < succ[k] ? key : succ[k];
      }
    }
  }
  const int col = col0 + threadIdx.x;
  double* out = partial + (long long)blockIdx.x * 7 * d;
  // sums
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int k = 0; k < BS_K; ++k) red[warp][lane + 32 * k] = q == 0 ? wlt[k] : (q == 1 ? weq[k] : neq[k]);
    __syncthreads();
    if (col < d) {
      double s = 0.0;
      for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
      out[(long long)q * d + col] = s;
    }
    __syncthreads();
  }
  // pred (max) / succ (min)
#pragma unroll
  for (int k = 0; k < BS_K; ++k) redk[warp][lane + 32 * k] = pred[k];
  __syncthreads();
  if (col < d) {
    unsigned long long s = 0ull;
    for (int wv = 0; wv < 8; ++wv) s = redk[wv][threadIdx.x] > s ? redk[wv][threadIdx.x] : s;
    out[(long long)3 * d + col] = __longlong_as_double((long long)s);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < BS_K; ++k) redk[warp][lane + 32 * k] = succ[k];
  __syncthreads();
  if (col < d) {
    unsigned long long s = ~0ull;
    for (int wv = 0; wv < 8; ++wv) s = redk[wv][threadIdx.x] < s ? redk[wv][threadIdx.x] : s;
    out[(long long)4 * d + col] = __longlong_as_double((long long)s);
  }
  __syncthreads();
  // first tie position and its weight
#pragma unroll
  for (int k = 0; k < BS_K; ++k) {
    redk[warp][lane + 32 * k] = fpos[k];
    red[warp][lane + 32 * k] = fw[k];
  }
  __syncthreads();
  if (col < d) {
    unsigned long long s = ~0ull;
    double fwv = 0.0;
    for (int wv = 0; wv < 8; ++wv)
      if (redk[wv][threadIdx.x] < s) {
        s = redk[wv][threadIdx.x];
        fwv = red[wv][threadIdx.x];
      }
    out[(long long)5 * d + col] = __longlong_as_double((long long)s);
    out[(long long)6 * d + col] = fwv;
  }
}

// Combine the tile statistics per (bin, col) -> stats[q][bin][col], q as listed in include/nls_b200.h.
__global__ void bs_stats_reduce_kernel(const double* __restrict__ partial, const int2* __restrict__ bin_tiles, int nbins,
                                       int d, const unsigned long long* __restrict__ vkey,
                                       double* __restrict__ stats) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nbins * d) return;
  const int b = e / d, col = e % d;
  double wlt = 0.0, weq = 0.0, neq = 0.0, fw = 0.0;
  unsigned long long pred = 0ull, succ = ~0ull, fpos = ~0ull;
  for (int t = bin_tiles[b].x; t < bin_tiles[b].y; ++t) {
    const double* p = partial + (long long)t * 7 * d;
    wlt += p[col];
    weq += p[(long long)d + col];
    neq += p[(long long)2 * d + col];
    const unsigned long long pk = (unsigned long long)__double_as_longlong(p[(long long)3 * d + col]);
    const unsigned long long sk = (unsigned long long)__double_as_longlong(p[(long long)4 * d + col]);
    const unsigned long long fk = (unsigned long long)__double_as_longlong(p[(long long)5 * d + col]);
    pred = pk > pred ? pk : pred;
    succ = sk < succ ? sk : succ;
    if (fk < fpos) {
      fpos = fk;
      fw = p[(long long)6 * d + col];
    }
  }
  const long long nd = (long long)nbins * d;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  stats[e] = keyd(vkey[e]);
  stats[nd + e] = pred == 0ull ? nan : keyd(pred);
  stats[2 * nd + e] = succ == ~0ull ? nan : keyd(succ);
  stats[3 * nd + e] = wlt;
  stats[4 * nd + e] = weq;
  stats[5 * nd + e] = neq;
  stats[6 * nd + e] = fw;
}

// partial[tile, col] = sum over the tile's rows of w |x - centre[bin, col]|
__global__ void __launch_bounds__(256) bs_mad_kernel(const double* __restrict__ X, int d,
                                                     const long long* __restrict__ perm, const double* __restrict__ w,
                                                     const BsTile* __restrict__ tiles,
                                                     const double* __restrict__ centre, double* __restrict__ partial) {
  __shared__ double red[8][BS_COLS];
  const BsTile t = tiles[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.y * BS_COLS;
  double c[BS_K], acc[BS_K];
#pragma unroll
  for (int k = 0; k < BS_K; ++k) {
    const int col = col0 + lane + 32 * k;
    c[k] = col < d ? centre[(long long)t.bin * d + col] : 0.0;
    acc[k] = 0.0;
  }
  for (int r = t.r0 + warp; r < t.r1; r += 8) {
    const double* x = X + perm[r] * d;
    const double wr = w[r];
#pragma unroll
    for (int k = 0; k < BS_K; ++k) {
      const int col = col0 + lane + 32 * k;
      if (col < d) acc[k] += wr * fabs(x[col] - c[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < BS_K; ++k) red[warp][lane + 32 * k] = acc[k];
  __syncthreads();
  const int col = col0 + threadIdx.x;
  if (col < d) {
    double s = 0.0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    partial[(long long)blockIdx.x * d + col] = s;
  }
}

__global__ void bs_sum_kernel(const double* __restrict__ partial, const int2* __restrict__ bin_tiles, int nbins, int d,
                              double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nbins * d) return;
  const int b = e / d, col = e % d;
  double F = 0.0;
  for (int t = bin_tiles[b].x; t < bin_tiles[b].y; ++t) F += partial[(long long)t * d + col];
  out[e] = F;
}

}  // namespace nls
