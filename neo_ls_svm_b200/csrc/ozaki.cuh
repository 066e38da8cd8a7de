// ozaki.cuh — FP64-accurate GEMM on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators) by the Ozaki scheme.
//
// sm_100a has no FP64 kind on tcgen05; DMMA tops out at 128 flop/cycle/SM.  The Gram (reference: _neo_ls_svm.py:112-114)
// and the eigenbasis projection T = phi Q (:134, :137) — 70 % of the C3 fit — have bounded operands (|phi| <= D^-1/2,
// Q unitary), so each operand row is scaled by a power of two to |x| <= 0.495, rounded ONCE to 56 fractional bits
// (more than FP64 carries) and cut exactly into S = 7 balanced base-256 digits x = 2^e sum_p d_p 256^-(p+1),
// d_p in [-128, 127].  A digit-plane product A_p B_q^T is an exact INT8 x INT8 -> INT32 GEMM (K <= 16384 per work
// item keeps every level inside INT32); the S(S+1)/2 = 28 products with p + q <= 6 go into 7 TMEM accumulators (one
// per level t = p + q, 7 x 64 = 448 of the 512 columns) and are recombined in FP64 in the epilogue, smallest level
// first.  Dropped levels (p + q >= 7) are below 2^-58 of max|a| max|b| per term.
//
// Data path (everything is laid out by our own slicing kernels, so no tensor maps are needed):
//   * planes live in global memory TILE-MAJOR: [row block][k step][plane][rows x 32 bytes], each [rows x 32 B] tile
//     already in the canonical K-major SWIZZLE_32B shared-memory image.  A pipeline stage (one k step of 32: 7 A
//     tiles of 128 rows + 7 B tiles of 64 rows = 42 KB) is two contiguous blocks, fetched with two 1-D
//     cp.async.bulk copies that complete on the stage's mbarrier.
//   * persistent CTAs (one per SM), warp-specialised: warp 0 = bulk-copy producer (one lane), warps 1-4 = MMA issuers
//     (one lane each, 7 of the 28 tcgen05.mma per stage, tcgen05.commit releases the stage; descriptors must stay in UNIFORM
//     registers — a 64-bit division feeding the loop bounds once turned every MMA into an ELECT + 4 R2UR.BROADCAST
//     sequence and cost 25 % of the rate, so work indices are 32-bit), warps 5-12 = epilogue (tcgen05.ld,
//     FP64 recombination, fused epilogue functor).  5 stages in flight; the accumulators are handed back to the
//     MMA warp as soon as they are drained into registers, so the epilogue math and stores of tile i overlap the
//     mainloop of tile i + 1.
//   * B tile rows 0..31 are the "real" rows and rows 32..63 the "imaginary" rows of the same 32 complex columns,
//     so one thread owns Re and Im of its (row, column) pairs.
#pragma once
#include "ptx.cuh"

namespace nls {
namespace oz {

constexpr int S = 7;               // digit planes per operand (signed 8-bit digits)
constexpr int TM = 128;            // rows of A per tile
constexpr int TN = 64;             // rows of B per tile (32 complex columns: 32 "re" rows then 32 "im" rows)
constexpr int KS = 32;             // k elements (bytes) per pipeline stage = one tcgen05.mma kind::i8 K step
constexpr int NSTAGE = 5;
constexpr int A_TILE = TM * KS;    // 4096
constexpr int B_TILE = TN * KS;    // 2048
constexpr int A_STAGE = S * A_TILE;
constexpr int B_STAGE = S * B_TILE;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;  // 43008
constexpr int SMEM_BYTES = 1024 + NSTAGE * STAGE_BYTES;
constexpr int DEFAULT_TMEM_PLANES = 7;  // leading A planes that go through tensor memory (A_TMEM kernels)
constexpr int NISSUE = 4;          // MMA-issuing warps: issuer w owns the levels {6}, {5,0}, {4,1}, {3,2} (7 products each)
constexpr int THREADS = 32 * (1 + NISSUE + 8);  // producer warp, MMA issuers, 8 epilogue warps
constexpr int RADIX_BITS = 8;      // balanced base-256 digits
constexpr int FRAC_BITS = RADIX_BITS * S;  // 56: more than the 53 bits of the FP64 operands
constexpr uint32_t TMEM_COLS = 512;
constexpr int IMAGE = 6;           // tile image the product uses: canonical K-major SWIZZLE_32B (see tile_off)
constexpr int MAX_K = 16384;       // INT32 accumulators: (t + 1) 128^2 K < 2^31 for every level t <= 6 (K per work item)

// ---- tile images ----------------------------------------------------------------------------------
// Byte offset of the 16-byte chunk c (0/1) of row r inside a [rows x 32 B] tile.
//   LAYOUT 6: canonical K-major SWIZZLE_32B (row pitch 32 B, chunk index XOR address bit 7)
//   LAYOUT 0: canonical K-major no-swizzle ("interleave": 8 x 16 B core matrices, LBO between the two k chunks)
template <int LAYOUT>
__host__ __device__ __forceinline__ uint32_t tile_off(int r, int c) {
  if (LAYOUT == 6) return (uint32_t)(r * 32 + ((c ^ ((r >> 2) & 1)) << 4));
  return (uint32_t)((r >> 3) * 256 + c * 128 + (r & 7) * 16);
}

template <int LAYOUT>
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(LAYOUT == 6 ? 1 : (128 >> 4)) << 16;  // leading byte offset (k chunk pitch; unused when swizzled)
  d |= (uint64_t)(256 >> 4) << 32;                      // stride byte offset: 8-row groups are 256 B apart
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)LAYOUT << 61;
  return d;
}

// D = S32, A = B = signed 8 bit, both K-major, M = 128, N = 64
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// ---- digit extraction -----------------------------------------------------------------------------
// x 2^-e in [-0.495, 0.495]  ->  V = rint(x 2^(56 - e))  =  sum_p d_p 256^(6 - p)  with BALANCED base-256 digits
// d_p in [-128, 127] (every plane is a signed int8).  With W = V + 0x80..80 (seven bytes), digit p is byte (6 - p)
// of W with its top bit flipped.  `scale` = 2^(56 - e); the product x * scale is exact, the conversion rounds once.
__host__ __device__ __forceinline__ long long quantise(double x, double scale) {
#ifdef __CUDA_ARCH__
  // saturate instead of wrapping if a caller ever breaks the |x| 2^-e <= 0.495 contract (e.g. a non-unitary basis handed to
  // the C ABI): the result is then inaccurate, not garbage
  constexpr long long LIM = (127LL << 48) | 0x7f7f7f7f7f7fLL;  // largest value the seven balanced digits represent
  return max(min(__double2ll_rn(x * scale), LIM), -LIM);
#else
  return (long long)__builtin_llrint(x * scale);
#endif
}
constexpr long long DIGIT_BIAS = 0x0080808080808080LL;
// all seven digits of one value as bytes (byte 6 - p = digit p, two's complement int8)
__host__ __device__ __forceinline__ unsigned long long digit_bytes(long long v) {
  return (unsigned long long)(v + DIGIT_BIAS) ^ (unsigned long long)DIGIT_BIAS;
}
__host__ __device__ __forceinline__ int digit(unsigned long long bytes, int p) {  // p = 0 is the leading digit
  return (int)(signed char)((bytes >> (8 * (S - 1 - p))) & 0xff);
}
// weight of level t = p + q in the recombination: 256^-(t + 2)
__host__ __device__ __forceinline__ double level_weight(int t) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)(1023 - RADIX_BITS * (t + 2)) << 52);
#else
  return __builtin_ldexp(1.0, -RADIX_BITS * (t + 2));
#endif
}

// Power-of-two exponent e with |x| 2^-e <= 0.495 for all |x| <= amax (0 when amax == 0): the balanced digits reach
// 127/255 (256^7 - 1)/256^7 = 0.498 at most.
__host__ __device__ __forceinline__ int scale_exponent(double amax) {
  if (!(amax > 0.0)) return 0;
  int ex;
  const double f = frexp(amax, &ex);  // amax = f 2^ex, f in [1/2, 1)
  return f <= 0.99 ? ex + 1 : ex + 2;
}

// ---- slicing kernels ------------------------------------------------------------------------------
// K order of the projection-type products (K = the feature index of both halves of phi): k step 2 fb holds the cos
// features 32 fb .. 32 fb + 31, k step 2 fb + 1 the sin features of the same block, nks = 2 ceil(D / 32).  With an
// upper-triangular B (predict_std through U^-1) column tile nb then needs the contiguous range ks < 2 (nb + 1) only.
__host__ __device__ __forceinline__ int feature_ksteps(int D) { return 2 * ((D + KS - 1) / KS); }

// A operand: `rows` rows of the planar feature chunk psi = [C | S] (pitch ld, each half Dp wide, D valid), one
// global exponent e.  Work item = one 16-byte chunk of one row of one k step; items run (chunk, row) fastest so
// that a warp writes 512 contiguous bytes per plane.
template <int LAYOUT>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ X, long long ld, int rows, int D, int Dp,
                                                         double scale, int nks, int row_blocks, int8_t* __restrict__ out) {
  const long long items = (long long)row_blocks * nks * TM * 2;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(it & 1);
    const int r = (int)((it >> 1) & (TM - 1));
    const long long blk = it >> 8;  // rb * nks + ks
    const int ks = (int)(blk % nks);
    const long long rb = blk / nks;
    const long long row = rb * TM + r;
    const int col0 = (ks >> 1) * KS + c * 16;  // first feature of this chunk, in half (ks & 1)
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[p][j] = 0;
    if (row < rows && col0 < D) {
      const double* src = X + row * ld + (ks & 1) * Dp + col0;
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const int col = col0 + j;
        double2 v = make_double2(0.0, 0.0);
        if (col < D) v = *reinterpret_cast<const double2*>(src + j);
        if (col + 1 >= D) v.y = 0.0;
        const unsigned long long q0 = digit_bytes(quantise(v.x, scale)), q1 = digit_bytes(quantise(v.y, scale));
#pragma unroll
        for (int p = 0; p < S; ++p) {
          const uint32_t b0 = (uint32_t)(q0 >> (8 * (S - 1 - p))) & 0xff, b1 = (uint32_t)(q1 >> (8 * (S - 1 - p))) & 0xff;
          w[p][j >> 2] |= (b0 | (b1 << 8)) << (16 * ((j >> 1) & 1));
        }
      }
    }
    int8_t* dst = out + (blk * S) * A_TILE + tile_off<LAYOUT>(r, c);
#pragma unroll
    for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(dst + (long long)p * A_TILE) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
  }
}

// B operand of the projection: Bt = [Re Q^T ; Im Q^T] (2 Np x Dp, as the DMMA path keeps it).  Complex column j
// becomes two rows of the real product over both halves of k:   re row = [ ya_j | yb_j ],  im row = [ -yb_j | ya_j ]
// (R = C ya + S yb = Re T,  I = S ya - C yb = -Im T: the convention of gemm_core's MODE_COMPLEX).
// Pass 1: per-column exponents and the recombination scale 2^(eA + eB_j).
__global__ void __launch_bounds__(256) basis_exponent_kernel(const double* __restrict__ Bt, int Np, int Dp, int D, int cols,
                                                             int cols_padded, int eA, int* __restrict__ ex,
                                                             double* __restrict__ colscale) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= cols_padded) return;
  double amax = 0.0;
  if (j < cols)
    for (int k = lane; k < D; k += 32) amax = fmax(amax, fmax(fabs(Bt[(long long)j * Dp + k]), fabs(Bt[(long long)(Np + j) * Dp + k])));
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) {
    const int e = scale_exponent(amax);
    ex[j] = e;
    colscale[j] = ldexp(1.0, eA + e);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) slice_basis_kernel(const double* __restrict__ Bt, int Np, int Dp, int D, int cols,
                                                          const int* __restrict__ ex, int nks, int n_tiles,
                                                          int8_t* __restrict__ out) {
  const long long items = (long long)n_tiles * nks * TN * 2;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(it & 1);
    const int r = (int)((it >> 1) & (TN - 1));
    const long long blk = it >> 7;  // nb * nks + ks
    const int ks = (int)(blk % nks);
    const int nb = (int)(blk / nks);
    const int j = nb * (TN / 2) + (r & (TN / 2 - 1));
    const bool im = r >= TN / 2;
    const int half = ks & 1, col0 = (ks >> 1) * KS + c * 16;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) w[p][q] = 0;
    if (j < cols && col0 < D) {
      // re row: half 0 -> ya, half 1 -> yb;   im row: half 0 -> -yb, half 1 -> ya
      const bool use_b = (half == 1) != im;
      const double sign = (im && half == 0) ? -1.0 : 1.0;
      const double* src = Bt + (long long)((use_b ? Np : 0) + j) * Dp + col0;
      const double scale = sign * ldexp(1.0, FRAC_BITS - ex[j]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const double v = col0 + i < D ? src[i] : 0.0;
        const unsigned long long q = digit_bytes(quantise(v, scale));
#pragma unroll
        for (int p = 0; p < S; ++p) w[p][i >> 2] |= ((uint32_t)(q >> (8 * (S - 1 - p))) & 0xff) << (8 * (i & 3));
      }
    }
    int8_t* dst = out + (blk * S) * B_TILE + tile_off<LAYOUT>(r, c);
#pragma unroll
    for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(dst + (long long)p * B_TILE) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
  }
}

// Gram operands from the weighted transposed feature chunk psiT (rows [0, DpT) = s_i cos, rows [DpT, 2 DpT) = s_i sin,
// pitch ldT, `rows` valid entries per row).  With k = the chunk's row index i, K = 2 Rp (Rp = rows rounded up to 32):
//   A side, feature f:            [ C_f(i) | S_f(i) ]
//   B side, re row of feature l:  [ C_l(i) | S_l(i) ]         im row:  [ -S_l(i) | C_l(i) ]
// so that R = C^T C + S^T S = Re A and I = S^T C - C^T S = Im A, the convention of OpGram.  One pass over psiT writes
// both operand images.  scale_dev[0] = 2^(56 - e) with e the exponent of this chunk's largest |entry| bound.
__global__ void __launch_bounds__(1024) gram_scale_kernel(const double* __restrict__ s, int rows, double dinv, double* __restrict__ out) {
  __shared__ double red[32];
  double amax = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) amax = fmax(amax, fabs(s[i]));
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) amax = fmax(amax, red[w]);
    const int e = scale_exponent(amax * dinv * (1.0 + 0x1p-50));  // fl(s cos / sqrt(D)) <= s_max D^-1/2 (1 + ulp)
    out[0] = ldexp(1.0, FRAC_BITS - e);
    out[1] = ldexp(1.0, 2 * e);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) slice_gram_kernel(const double* __restrict__ psiT, long long ldT, int DpT, int D, int rows,
                                                         const double* __restrict__ scale_dev, int nksR, int row_blocks,
                                                         int n_tiles, int8_t* __restrict__ GA, int8_t* __restrict__ GB) {
  const double scale = scale_dev[0];
  const long long items = (long long)row_blocks * 2 * nksR * TM * 2;
  const int nks2 = 2 * nksR;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(it & 1);
    const int r = (int)((it >> 1) & (TM - 1));
    long long rest = it >> 8;
    const int ks = (int)(rest % nksR);
    rest /= nksR;
    const int h = (int)(rest & 1);
    const int kb = (int)(rest >> 1);
    const int f = kb * TM + r;
    const int i0 = ks * KS + c * 16;
    uint32_t w[S][4], wn[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[p][j] = wn[p][j] = 0;
    if (f < D) {
      const double* src = psiT + ((long long)h * DpT + f) * ldT + i0;
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        double2 v = make_double2(0.0, 0.0);
        if (i0 + j < rows) v = *reinterpret_cast<const double2*>(src + j);
        if (i0 + j + 1 >= rows) v.y = 0.0;
        const long long v0 = quantise(v.x, scale), v1 = quantise(v.y, scale);
        const unsigned long long q0 = digit_bytes(v0), q1 = digit_bytes(v1), n0 = digit_bytes(-v0), n1 = digit_bytes(-v1);
#pragma unroll
        for (int p = 0; p < S; ++p) {
          const int sh = 8 * (S - 1 - p);
          w[p][j >> 2] |= (((uint32_t)(q0 >> sh) & 0xff) | (((uint32_t)(q1 >> sh) & 0xff) << 8)) << (16 * ((j >> 1) & 1));
          if (h) wn[p][j >> 2] |= (((uint32_t)(n0 >> sh) & 0xff) | (((uint32_t)(n1 >> sh) & 0xff) << 8)) << (16 * ((j >> 1) & 1));
        }
      }
    }
    int8_t* dst = GA + ((long long)kb * nks2 + h * nksR + ks) * S * A_TILE + tile_off<LAYOUT>(r, c);
#pragma unroll
    for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(dst + (long long)p * A_TILE) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
    const int lb = f >> 5, r32 = f & 31;
    if (lb < n_tiles) {
      // re row: same K position and digits as the A side
      dst = GB + ((long long)lb * nks2 + h * nksR + ks) * S * B_TILE + tile_off<LAYOUT>(r32, c);
#pragma unroll
      for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(dst + (long long)p * B_TILE) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
      // im row: C goes to the second K half as it is, S to the first half negated
      dst = GB + ((long long)lb * nks2 + (h ? 0 : nksR) + ks) * S * B_TILE + tile_off<LAYOUT>(32 + r32, c);
#pragma unroll
      for (int p = 0; p < S; ++p)
        *reinterpret_cast<uint4*>(dst + (long long)p * B_TILE) =
            h ? make_uint4(wn[p][0], wn[p][1], wn[p][2], wn[p][3]) : make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
    }
  }
}

// ---- γ sweep operands --------------------------------------------------------------------------------
// Column scaling per γ group: gref = the smallest γ of the group's 128-γ blocks, d[grp][k] = |λ_k| + gref.
__global__ void sweep_groups_kernel(const double* __restrict__ gammas, const double* __restrict__ lam, int G, int m, int ngrp,
                                    int rb_per_group, long long ldk, double* __restrict__ d, double* __restrict__ inv_d) {
  const int grp = blockIdx.x;
  const int g0 = grp * rb_per_group * TM, g1 = min(G, g0 + rb_per_group * TM);
  double gref = gammas[g0];
  for (int g = g0 + 1; g < g1; ++g) gref = fmin(gref, gammas[g]);
  for (int k = threadIdx.x; k < ldk; k += blockDim.x) {
    const double dk = k < m ? fabs(lam[k]) + gref : 1.0;
    d[grp * ldk + k] = dk;
    inv_d[grp * ldk + k] = 1.0 / dk;
  }
}

// r'_kg = d_k / (γ_g + λ_k): one warp per γ finds the row exponent; ascale[g] = 2^e.
__global__ void __launch_bounds__(256) rgamma_exponent_kernel(const double* __restrict__ gammas, const double* __restrict__ lam,
                                                              const double* __restrict__ d, int G, int m, int rb_per_group,
                                                              long long ldk, int* __restrict__ ex, double* __restrict__ ascale) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= G) return;
  const double* dg = d + (long long)(g / (rb_per_group * TM)) * ldk;
  const double gam = gammas[g];
  double amax = 0.0;
  for (int k = lane; k < m; k += 32) amax = fmax(amax, fabs(dg[k] / (gam + lam[k])));
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) {
    const int e = scale_exponent(amax);
    ex[g] = e;
    ascale[g] = ldexp(1.0, e);
  }
}

// A side: digit planes of r' [γ block of 128][k step][plane][128 x 32 B]
template <int LAYOUT>
__global__ void __launch_bounds__(256) slice_rgamma_kernel(const double* __restrict__ gammas, const double* __restrict__ lam,
                                                           const double* __restrict__ d, const int* __restrict__ ex, int G, int m,
                                                           int rb_per_group, long long ldk, int nks, int row_blocks,
                                                           int8_t* __restrict__ out) {
  const long long items = (long long)row_blocks * nks * TM * 2;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(it & 1);
    const int r = (int)((it >> 1) & (TM - 1));
    const long long blk = it >> 8;
    const int ks = (int)(blk % nks);
    const int gb = (int)(blk / nks);
    const int g = gb * TM + r;
    const int k0 = ks * KS + c * 16;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[p][j] = 0;
    if (g < G && k0 < m) {
      const double* dg = d + (long long)(gb / rb_per_group) * ldk;
      const double gam = gammas[g];
      const double scale = ldexp(1.0, FRAC_BITS - ex[g]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = k0 + i;
        const double v = k < m ? dg[k] / (gam + lam[k]) : 0.0;
        const unsigned long long q = digit_bytes(quantise(v, scale));
#pragma unroll
        for (int p = 0; p < S; ++p) w[p][i >> 2] |= ((uint32_t)(q >> (8 * (S - 1 - p))) & 0xff) << (8 * (i & 3));
      }
    }
    int8_t* dst = out + (blk * S) * A_TILE + tile_off<LAYOUT>(r, c);
#pragma unroll
    for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(dst + (long long)p * A_TILE) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
  }
}

// B side, pass 1: row exponents of P' = P / d and U' = U / d per (data row, group); one warp per data row.
constexpr int MAX_SWEEP_GROUPS = 4;
__global__ void __launch_bounds__(256) pu_exponent_kernel(const double* __restrict__ P, const double* __restrict__ U, long long ld, int rows,
                                                          int m, const double* __restrict__ inv_d, int ngrp, long long ldk,
                                                          int* __restrict__ ex /*[rows][ngrp][2]*/, double* __restrict__ pscale,
                                                          double* __restrict__ uscale) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= rows) return;
  double pm[MAX_SWEEP_GROUPS], um[MAX_SWEEP_GROUPS];
#pragma unroll
  for (int grp = 0; grp < MAX_SWEEP_GROUPS; ++grp) pm[grp] = um[grp] = 0.0;
  // the row is read once (16-byte loads; ld and ldk are multiples of 16), every group's scaling applied on the fly
  for (int k = 2 * lane; k < m; k += 64) {
    const double2 pv = *reinterpret_cast<const double2*>(P + (long long)i * ld + k);
    const double2 uv = *reinterpret_cast<const double2*>(U + (long long)i * ld + k);
    const bool two = k + 1 < m;
#pragma unroll
    for (int grp = 0; grp < MAX_SWEEP_GROUPS; ++grp)
      if (grp < ngrp) {
        const double2 id = *reinterpret_cast<const double2*>(inv_d + grp * ldk + k);
        pm[grp] = fmax(pm[grp], fmax(fabs(pv.x * id.x), two ? fabs(pv.y * id.y) : 0.0));
        um[grp] = fmax(um[grp], fmax(fabs(uv.x * id.x), two ? fabs(uv.y * id.y) : 0.0));
      }
  }
#pragma unroll
  for (int grp = 0; grp < MAX_SWEEP_GROUPS; ++grp)
    if (grp < ngrp) {
      double a = pm[grp], b = um[grp];
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
      }
      if (lane == 0) {
        const int ep = scale_exponent(a), eu = scale_exponent(b);
        ex[((long long)i * ngrp + grp) * 2] = ep;
        ex[((long long)i * ngrp + grp) * 2 + 1] = eu;
        pscale[(long long)i * ngrp + grp] = ldexp(1.0, ep);
        uscale[(long long)i * ngrp + grp] = ldexp(1.0, eu);
      }
    }
}

// Global exponents of P' and U' per γ group from the bounds of EpiProjectPlanes: qscale[grp][2] = 2^(56 - e), gscale = 2^e.
__global__ void __launch_bounds__(256) sweep_bounds_kernel(const double* __restrict__ v_r, const double* __restrict__ v_i,
                                                           const double* __restrict__ inv_d, int m, long long ldk, double inv_c,
                                                           double* __restrict__ qscale, double* __restrict__ gscale) {
  __shared__ double red[2][8];
  const int grp = blockIdx.x;
  double bp = 0.0, bu = 0.0;
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const double id = inv_d[grp * ldk + k];
    bp = fmax(bp, hypot(v_r[k], v_i[k]) * id);
    bu = fmax(bu, id);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    bp = fmax(bp, __shfl_xor_sync(0xffffffffu, bp, o));
    bu = fmax(bu, __shfl_xor_sync(0xffffffffu, bu, o));
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = bp;
    red[1][threadIdx.x >> 5] = bu;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      bp = fmax(bp, red[0][w]);
      bu = fmax(bu, red[1][w]);
    }
    // |T| <= |phi| |q| = sqrt(2) up to rounding, |T|^2 <= 2
    const int ep = scale_exponent(1.41422 * bp), eu = scale_exponent(2.0001 * inv_c * bu);
    qscale[2 * grp] = ldexp(1.0, FRAC_BITS - ep);
    qscale[2 * grp + 1] = ldexp(1.0, FRAC_BITS - eu);
    gscale[2 * grp] = ldexp(1.0, ep);
    gscale[2 * grp + 1] = ldexp(1.0, eu);
  }
}

// B side, pass 2: digit planes [tile of 32 data rows][group][k step][plane][64 x 32 B]; tile rows 0..31 = P', 32..63 = U'.
// k steps [ks_begin, nks) only (the spill columns when the projection epilogue wrote the others); with qscale != null
// the global scales of sweep_bounds_kernel are used instead of the row exponents.
template <int LAYOUT>
__global__ void __launch_bounds__(256) slice_pu_kernel(const double* __restrict__ P, const double* __restrict__ U, long long ld, int rows,
                                                       int m, const double* __restrict__ inv_d, const int* __restrict__ ex,
                                                       const double* __restrict__ qscale, int ngrp, long long ldk, int ks_begin,
                                                       int nks, int n_tiles, int8_t* __restrict__ out) {
  const int nk = nks - ks_begin;
  const long long items = (long long)n_tiles * ngrp * nk * TN * 2;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(it & 1);
    const int r = (int)((it >> 1) & (TN - 1));
    const long long sub = it >> 7;  // (nb * ngrp + grp) * nk + (ks - ks_begin)
    const int ks = ks_begin + (int)(sub % nk);
    const long long tg = sub / nk;
    const long long blk = tg * nks + ks;
    const int grp = (int)(tg % ngrp);
    const long long nb = tg / ngrp;
    const long long i = nb * (TN / 2) + (r & (TN / 2 - 1));
    const int which = r >> 5;  // 0: P', 1: U'
    const int k0 = ks * KS + c * 16;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[p][j] = 0;
    if (i < rows && k0 < m) {
      const double* src = (which ? U : P) + i * ld + k0;
      const double* id = inv_d + grp * ldk + k0;
      const double scale = qscale ? qscale[2 * grp + which] : ldexp(1.0, FRAC_BITS - ex[(i * ngrp + grp) * 2 + which]);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + j);  // pads of the P / U rows are never used: masked below
        const double2 dd = *reinterpret_cast<const double2*>(id + j);
        const unsigned long long q0 = digit_bytes(quantise(k0 + j < m ? v.x * dd.x : 0.0, scale));
        const unsigned long long q1 = digit_bytes(quantise(k0 + j + 1 < m ? v.y * dd.y : 0.0, scale));
#pragma unroll
        for (int p = 0; p < S; ++p) {
          const uint32_t b0 = (uint32_t)(q0 >> (8 * (S - 1 - p))) & 0xff, b1 = (uint32_t)(q1 >> (8 * (S - 1 - p))) & 0xff;
          w[p][j >> 2] |= (b0 | (b1 << 8)) << (16 * ((j >> 1) & 1));
        }
      }
    }
    int8_t* dst = out + (blk * S) * B_TILE + tile_off<LAYOUT>(r, c);
#pragma unroll
    for (int p = 0; p < S; ++p) *reinterpret_cast<uint4*>(dst + (long long)p * B_TILE) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
  }
}

// ---- pipeline primitives --------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
               : "memory");
}
// A operand from tensor memory (128 lanes x 8 columns per plane and k step), B from shared memory
__device__ __forceinline__ void umma_i8_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, 1, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(IDESC)
               : "memory");
}
// one [128 rows x 32 B] tile from shared memory into 8 TMEM columns (row r -> lane r)
__device__ __forceinline__ void tmem_cp_tile(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
               "r"(0u)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// exact int32 -> double without the conversion unit: 2^52 + 2^31 + x has x + 2^31 in its low mantissa word
__device__ __forceinline__ double i2d(uint32_t x) { return __hiloint2double(0x43300000, (int)(x ^ 0x80000000u)) - 4503601774854144.0; }

struct GemmParams {
  const int8_t* A;   // [row_blocks][nks][S][128 x 32 B]
  const int8_t* B;   // [n_tiles][nks][S][64 x 32 B]
  int nks, row_blocks, n_tiles;
  int upper;         // 1: only the tiles that touch the upper triangle (nb >= 4 rb): Hermitian Gram; 0: all of them
  int tiles;         // tiles per K split (row_blocks * n_tiles, or the upper-triangular count)
  int splits;        // the K range is cut into `splits` work items per tile (each with its own epilogue slot) ...
  int ks_per_split;  // ... of this many k steps
  int upper_k;       // 1: B is upper triangular in (feature, column): column tile nb needs ks < 2 (nb + 1) only
  int b_groups;      // B holds b_groups images per column tile; row block rb uses image rb / rb_per_group (γ sweep) ...
  int rb_per_group;  // ... (1 image: b_groups = 1, rb_per_group >= row_blocks)
  int rb_fastest;    // 1: work items run column tile by column tile with the row block fastest (few row blocks, many tiles)
};

struct Work {
  int rb, nb, ks0, ks1, split;
};

// Work items run split-major, row block next, column tile fastest: CTAs that run side by side share operand blocks in L2.
__device__ __forceinline__ Work get_work(const GemmParams& g, int w) {
  Work k;
  k.split = w / g.tiles;
  int idx = w - k.split * g.tiles;
  if (g.rb_fastest) {
    k.nb = idx / g.row_blocks;
    k.rb = idx - k.nb * g.row_blocks;
  } else if (!g.upper) {
    k.rb = idx / g.n_tiles;
    k.nb = idx % g.n_tiles;
  } else {
    int kb = 0;
    while (idx >= g.n_tiles - 4 * kb) {
      idx -= g.n_tiles - 4 * kb;
      ++kb;
    }
    k.rb = kb;
    k.nb = 4 * kb + idx;
  }
  k.ks0 = k.split * g.ks_per_split;
  k.ks1 = min(g.nks, k.ks0 + g.ks_per_split);
  if (g.upper_k) k.ks1 = min(k.ks1, 2 * (k.nb + 1));
  return k;
}

// Epilogue functors receive, per thread, one row and 16 consecutive complex columns: sr/si = the recombined real /
// imaginary sums BEFORE the column scale 2^(eA + eB_j).
struct EpiProject {  // P = Re(T v), U = |T|^2 / c   (what OpProject writes; reference _neo_ls_svm.py:134, :137)
  struct Params {
    int n_rows, m;  // valid rows, valid complex columns
    const double* colscale;
    const double* bias_r;
    const double* bias_i;
    const double* v_r;
    const double* v_i;
    double inv_c;
    double* P;
    double* U;
    long long ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, const Work&, long long row, int col0, const double (&sr)[16],
                                               const double (&si)[16]) {
    if (row >= p.n_rows) return;
    double pv[16], uv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = col0 + j;
      const bool ok = col < p.m;
      const double cs = ok ? p.colscale[col] : 0.0;
      const double tr = sr[j] * cs + (ok ? p.bias_r[col] : 0.0);
      const double ti = (ok ? p.bias_i[col] : 0.0) - si[j] * cs;
      pv[j] = tr * (ok ? p.v_r[col] : 0.0) - ti * (ok ? p.v_i[col] : 0.0);
      uv[j] = (tr * tr + ti * ti) * p.inv_c;
    }
    double* P = p.P + row * p.ld + col0;
    double* U = p.U + row * p.ld + col0;
    if (col0 + 16 <= p.m) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        *reinterpret_cast<double2*>(P + j) = make_double2(pv[j], pv[j + 1]);
        *reinterpret_cast<double2*>(U + j) = make_double2(uv[j], uv[j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col0 + j < p.m) {
          P[j] = pv[j];
          U[j] = uv[j];
        }
    }
  }
};

// The projection feeding the INT8 sweep directly: instead of P and U in FP64 it writes the digit planes of the sweep's B
// operand, P' = P / d and U' = U / d for every γ group.  A thread owns one data row and 16 consecutive columns k — exactly
// one 16-byte chunk of one k step of the sweep — so every (group, operand, plane) is one 16-byte store and the FP64 round
// trip (write P, U; row maxima; read twice; write planes) disappears.  The exponents cannot come from row maxima here (a
// row is spread over all column tiles), so they come from BOUNDS: |T_ik| <= |phi_i| |q_k| = sqrt(2), hence
// |P'_ik| <= sqrt(2) |v_k| / d_k and U'_ik <= 2 / (c d_k), maximised over k per group (sweep_bounds_kernel).  Typical entries are
// sqrt(m) (P') and m (U') below the bounds, i.e. 51 and 46 of the 56 bits stay significant: still FP64-level for the sums.
struct EpiProjectPlanes {
  struct Params {
    int n_rows, m;  // valid rows, valid complex columns (columns >= m are written as zeros)
    const double* colscale;
    const double* bias_r;
    const double* bias_i;
    const double* v_r;
    const double* v_i;
    double inv_c;
    const double* inv_d;   // [ngrp][ldk]
    const double* qscale;  // [ngrp][2] 2^(56 - e) of P' / U'
    int ngrp, nks;         // groups, k steps of the sweep (per (row tile, group): nks stages of 7 x [64 x 32 B])
    long long ldk;
    int8_t* B;             // [row tile of 32][group][k step][plane][64 x 32 B]
  };
  static __device__ __forceinline__ void apply(const Params& p, const Work&, long long row, int col0, const double (&sr)[16],
                                               const double (&si)[16]) {
    double pv[16], uv[16];
    const bool live = row < p.n_rows;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = col0 + j;
      const bool ok = live && col < p.m;
      const double cs = ok ? p.colscale[col] : 0.0;
      const double tr = sr[j] * cs + (ok ? p.bias_r[col] : 0.0);
      const double ti = (ok ? p.bias_i[col] : 0.0) - si[j] * cs;
      pv[j] = ok ? tr * p.v_r[col] - ti * p.v_i[col] : 0.0;
      uv[j] = ok ? (tr * tr + ti * ti) * p.inv_c : 0.0;
    }
    const long long tb = row >> 5;
    const int r32 = (int)(row & 31), ks = col0 >> 5, c = (col0 >> 4) & 1;
    for (int grp = 0; grp < p.ngrp; ++grp) {
      const double* id = p.inv_d + grp * p.ldk + col0;
      int8_t* base = p.B + (((tb * p.ngrp + grp) * p.nks + ks) * S) * (long long)B_TILE;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const double scale = p.qscale[2 * grp + which];
        uint32_t w[S][4];
#pragma unroll
        for (int q = 0; q < S; ++q)
#pragma unroll
          for (int x = 0; x < 4; ++x) w[q][x] = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const unsigned long long d = digit_bytes(quantise((which ? uv[j] : pv[j]) * id[j], scale));
#pragma unroll
          for (int q = 0; q < S; ++q) w[q][j >> 2] |= ((uint32_t)(d >> (8 * (S - 1 - q))) & 0xff) << (8 * (j & 3));
        }
        int8_t* dst = base + tile_off<IMAGE>(r32 + 32 * which, c);
#pragma unroll
        for (int q = 0; q < S; ++q) *reinterpret_cast<uint4*>(dst + (long long)q * B_TILE) = make_uint4(w[q][0], w[q][1], w[q][2], w[q][3]);
      }
    }
  }
};

struct EpiStore {  // raw T planes (probe / tests): Tr[row][col] = sr * colscale, Ti likewise
  struct Params {
    int n_rows, m;
    const double* colscale;
    double* Tr;
    double* Ti;
    long long ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, const Work&, long long row, int col0, const double (&sr)[16],
                                               const double (&si)[16]) {
    if (row >= p.n_rows) return;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < p.m) {
        p.Tr[row * p.ld + col0 + j] = sr[j] * p.colscale[col0 + j];
        p.Ti[row * p.ld + col0 + j] = si[j] * p.colscale[col0 + j];
      }
  }
};

// Stage 4c / 5b: sigma2_i = sum_k |(phi B)_ik|^2 w_k (what OpVariance computes; reference _neo_ls_svm.py:184, :467-469).
// Every thread sums its 16 columns; part[2 nb + column half][row], reduced in fixed order by rowsum_reduce_kernel.
struct EpiVariance {
  struct Params {
    int n_rows, m;
    const double* colscale;
    const double* bias_r;
    const double* bias_i;
    const double* w;
    double* part;
    long long part_ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, const Work& wk, long long row, int col0, const double (&sr)[16],
                                               const double (&si)[16]) {
    if (row >= p.n_rows) return;
    double rs = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = col0 + j;
      if (col < p.m) {
        const double cs = p.colscale[col];
        const double tr = sr[j] * cs + p.bias_r[col];
        const double ti = p.bias_i[col] - si[j] * cs;
        rs += (tr * tr + ti * ti) * p.w[col];
      }
    }
    p.part[(long long)(2 * wk.nb + ((col0 >> 4) & 1)) * p.part_ld + row] = rs;
  }
};

// Stage 2 on the INT8 core: rows = features k (A side), columns = features l (B side), K = the chunk's rows.  Every
// (split, tile) work item owns its part of the partial-sum workspace [split][2][D][D] and accumulates into it chunk
// after chunk in a fixed order, exactly like OpGram (upper triangle only; gram_assemble_kernel mirrors it).
//   reference: _neo_ls_svm.py:112-114.
struct EpiGram {
  struct Params {
    int D;
    const double* scale2;  // device scalar 2^(2 e) of this chunk (e = exponent of max_i s_i D^-1/2)
    double* ws;
  };
  static __device__ __forceinline__ void apply(const Params& p, const Work& wk, long long row, int col0, const double (&sr)[16],
                                               const double (&si)[16]) {
    if (row >= p.D) return;
    const double cs = *p.scale2;
    double* wr = p.ws + (long long)wk.split * 2 * p.D * p.D + row * p.D;
    double* wi = wr + (long long)p.D * p.D;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int l = col0 + j;
      if (l < p.D && l >= row) {
        wr[l] += sr[j] * cs;
        wi[l] += si[j] * cs;
      }
    }
  }
};

// Stage 4b on the INT8 core.  Roles are transposed with respect to the DMMA sweep: the 128 tile rows (TMEM lanes) are
// γ values, the B tile holds 32 data rows twice — rows 0..31 the scaled numerator operand P', rows 32..63 the scaled
// leverage operand U' — so a thread owns num and den of its (γ, data row) pairs:
//   num_ig = sum_k P'_ik r'_kg,  den_ig = sum_k U'_ik r'_kg,   P' = P / d,  U' = U / d,  r'_kg = d_k / (γ_g + λ_k),
// d_k = |λ_k| + γ_ref(group of γ_g).  The two-sided scaling is what makes a fixed-point split FP64-accurate here: r'
// lies in (γ_ref / γ_g, 1] and P'_ik is the k-th term of the prediction at γ_ref, so both operands are well scaled
// along k (scripts/sweep_int8_study.py; the plain split loses 1e-10 on ill-conditioned fits).  Fused: LOO residual,
// classifier clip, |.|, s-weighted partial sums over the thread's 16 data rows (fixed order; sweep_reduce_kernel adds
// the (tile, half) partials in order), and the σ² stash.   reference: _neo_ls_svm.py:147-161.
struct EpiSweep {
  struct Params {
    int n_rows, G, ngrp, rb_per_group;
    const double* ascale;  // [G] 2^e of the r' row of every γ
    const double* pscale;  // [rows][ngrp] 2^e of the P' row
    const double* uscale;  // [rows][ngrp] 2^e of the U' row
    const double* gscale;  // when not null: [ngrp][2] global 2^e of P' / U' (planes written by EpiProjectPlanes), rows ignored
    const double* y;
    const double* s;
    int is_classifier;
    double* part;          // [2 n_tiles][3][G]
    double* den_out;       // optional σ² stash (U rγ), row pitch den_ld
    long long den_ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, const Work& wk, long long row, int col0, const double (&sr)[16],
                                               const double (&si)[16]) {
    const int g = (int)row;
    if (g >= p.G) return;
    const int grp = wk.rb / p.rb_per_group;
    const double as = p.ascale[g];
    const double gp = p.gscale ? as * p.gscale[2 * grp] : 0.0, gu = p.gscale ? as * p.gscale[2 * grp + 1] : 0.0;
    double e_abs = 0.0, e_cnt = 0.0, e_hng = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = col0 + j;
      if (i < p.n_rows) {
        const double num = sr[j] * (p.gscale ? gp : as * p.pscale[(long long)i * p.ngrp + grp]);
        const double den = si[j] * (p.gscale ? gu : as * p.uscale[(long long)i * p.ngrp + grp]);
        if (p.den_out) p.den_out[(long long)i * p.den_ld + g] = den;
        const double yi = p.y[i], wi = p.s[i];
        double loo = (num - yi) / (1.0 - wi * wi * den);
        if (p.is_classifier && ((yi > 0.0 && loo > 0.0) || (yi < 0.0 && loo < 0.0))) loo = 0.0;
        const double a = fabs(loo);
        e_abs += wi * a;
        if (p.is_classifier) {
          e_cnt += (a >= 1.0) ? wi : 0.0;
          e_hng += wi * fmax(0.0, a - 1.0);
        }
      }
    }
    double* o = p.part + (long long)(2 * wk.nb + ((col0 >> 4) & 1)) * 3 * p.G + g;
    o[0] = e_abs;
    o[p.G] = e_cnt;
    o[2 * p.G] = e_hng;
  }
};

// A_TMEM: the MMAs take their A operand from tensor memory.  At M = 128, N = 64 with both operands in shared memory a
// k step moves 210 KB through the SM's 128 B/clk shared-memory port (28 x 6 KB of operand reads + the 42 KB the bulk
// copies write) — the measured bound of the plain variant.  Copying each A plane once per k step into TMEM
// (tcgen05.cp, 8 columns per plane: 448 + 56 = 504 of the 512 columns) cuts that to 126 KB.  Issuer w then owns the A
// PLANES {0}, {1,6}, {2,5}, {3,4} (7 products each) instead of whole levels: its tcgen05.cp and the MMAs that read the
// plane stay in one thread's program order, which is the only ordering tcgen05 guarantees.  Levels are then shared
// between issuers, so no product may overwrite: the epilogue warps zero the accumulators (tcgen05.st) after draining.
template <int LAYOUT, class Epi, bool A_TMEM = true, int TMEM_PLANES = DEFAULT_TMEM_PLANES>
__global__ void __launch_bounds__(THREADS, 1) gemm_kernel_i8(GemmParams g, typename Epi::Params ep) {
  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 32) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), NISSUE);  // one tcgen05.commit per issuer
    }
    mbar_init(smem_u32(&tmem_full_bar), NISSUE);
    mbar_init(smem_u32(&tmem_empty_bar), 8);  // one arrival per epilogue warp
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t smem_base = smem_u32(smem);
  const int total_work = g.tiles * g.splits;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const Work wk = get_work(g, w);
        const int8_t* a = g.A + (long long)wk.rb * g.nks * (long long)A_STAGE;
        const int8_t* b = g.B + ((long long)wk.nb * g.b_groups + wk.rb / g.rb_per_group) * g.nks * (long long)B_STAGE;
        for (int ks = wk.ks0; ks < wk.ks1; ++ks, ++it) {
          const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_arrive_expect_tx(fb, STAGE_BYTES);
          bulk_g2s(smem_base + s * STAGE_BYTES, a + (long long)ks * A_STAGE, A_STAGE, fb);
          bulk_g2s(smem_base + s * STAGE_BYTES + A_STAGE, b + (long long)ks * B_STAGE, B_STAGE, fb);
        }
      }
    }
  } else if (warp <= NISSUE) {
    // One issuing lane costs ~60 cycles per tcgen05.mma (ELECT / R2UR / UTCIMMA / branch chain) against the 48 the
    // tensor pipe needs at this tile shape, so the 28 products are spread over NISSUE issuers.  An issuer owns WHOLE
    // levels: the products of one accumulator then stay in one thread's program order, which is what makes the
    // "first product of the tile overwrites" flag safe.
    if (lane == 0) {
      const int w = warp - 1;
      const int t_hi = 6 - w, t_lo = w == 0 ? -1 : w - 1;  // levels {6}, {5,0}, {4,1}, {3,2}
      const int p_a = w, p_b = w == 0 ? -1 : 7 - w;          // A planes {0}, {1,6}, {2,5}, {3,4}
      uint32_t it = 0, tile_i = 0;
      for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x, ++tile_i) {
        const Work wk = get_work(g, wi);
        // accumulators drained (A_TMEM: and zeroed, including once before the first tile) by the epilogue warps
        mbar_wait(smem_u32(&tmem_empty_bar), A_TMEM ? (tile_i & 1) : ((tile_i & 1) ^ 1));
        tc_fence_after();
        for (int ks = wk.ks0; ks < wk.ks1; ++ks, ++it) {
          const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint64_t da = smem_desc<LAYOUT>(smem_base + s * STAGE_BYTES);
          const uint64_t db = smem_desc<LAYOUT>(smem_base + s * STAGE_BYTES + A_STAGE);
          if (A_TMEM) {
#pragma unroll
            for (int p = 0; p < S; ++p) {
              if (p != p_a && p != p_b) continue;
              // a plane feeds 7 - p products: the copy (64 cycles of the tensor pipe) pays off against the 16 cycles it
              // saves per product only for the leading planes; the others are read from shared memory directly
              if (p < TMEM_PLANES) {
                tmem_cp_tile(tmem + (uint32_t)(S * TN + p * 8), da + (uint64_t)((p * A_TILE) >> 4));
#pragma unroll
                for (int q = 0; q < S; ++q)
                  if (p + q < S) umma_i8_ta(tmem + (uint32_t)((p + q) * TN), tmem + (uint32_t)(S * TN + p * 8), db + (uint64_t)((q * B_TILE) >> 4));
              } else {
#pragma unroll
                for (int q = 0; q < S; ++q)
                  if (p + q < S) umma_i8(tmem + (uint32_t)((p + q) * TN), da + (uint64_t)((p * A_TILE) >> 4), db + (uint64_t)((q * B_TILE) >> 4), 1u);
              }
            }
          } else {
          const uint32_t acc0 = ks > wk.ks0 ? 1u : 0u;
#pragma unroll
          for (int t = S - 1; t >= 0; --t) {
            if (t != t_hi && t != t_lo) continue;
#pragma unroll
            for (int p = 0; p < S; ++p)
              if (p <= t)  // level t = p + q; its first product of the tile (p = 0) overwrites, the rest accumulate
                umma_i8(tmem + (uint32_t)(t * TN), da + (uint64_t)((p * A_TILE) >> 4), db + (uint64_t)(((t - p) * B_TILE) >> 4),
                        p == 0 ? acc0 : 1u);
          }
          }
          umma_commit(smem_u32(&empty_bar[s]));  // stage reusable once every issuer's MMAs have read it
        }
        umma_commit(smem_u32(&tmem_full_bar));
      }
    }
  } else {
    const int e = warp - 1 - NISSUE;
    const int quad = warp & 3;       // the TMEM lane quadrant this warp may access
    const int cb = (e >> 2) * 16;    // first of its 16 complex columns within the tile
    uint32_t tile_i = 0;
    if (A_TMEM) {  // the first tile also starts from zeroed accumulators (fresh TMEM holds garbage)
#pragma unroll
      for (int t = 0; t < S; ++t) {
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * TN + cb);
        tmem_st16_zero(taddr);
        tmem_st16_zero(taddr + TN / 2);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar));
    }
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tile_i) {
      const Work wk = get_work(g, w);
      mbar_wait(smem_u32(&tmem_full_bar), tile_i & 1);
      tc_fence_after();
      double sr[16], si[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) sr[j] = si[j] = 0.0;
#pragma unroll
      for (int t = S - 1; t >= 0; --t) {
        uint32_t ar[16], ai[16];
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * TN + cb);
        tmem_ld16(taddr, ar);
        tmem_ld16(taddr + TN / 2, ai);
        tmem_ld_wait();
        if (A_TMEM) {  // hand the accumulators back zeroed: every product of the next tile accumulates
          tmem_st16_zero(taddr);
          tmem_st16_zero(taddr + TN / 2);
        }
        const double w = level_weight(t);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sr[j] = fma(i2d(ar[j]), w, sr[j]);
          si[j] = fma(i2d(ai[j]), w, si[j]);
        }
      }
      if (A_TMEM) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar));
      Epi::apply(ep, wk, (long long)wk.rb * TM + quad * 32 + lane, wk.nb * (TN / 2) + cb, sr, si);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

// ---- INT8 tensor peak of this device (bench.py's roofline denominator) -----------------------------
// Every CTA (one per SM) issues `iters` trains of four resident-tile MMAs (M = 128, N columns, K = 32 each; operand
// tiles in the K-major SWIZZLE_128B image, contents irrelevant) into one TMEM accumulator: no loads, no epilogue.
// N = 256 is the widest single-CTA shape (the hardware peak, 8192 MAC/cycle/SM); N = 64 is the shape the 7-level
// Ozaki tile has to use (7 x 64 TMEM columns), where every MMA is bound by its 6 KB of shared-memory operand reads.
template <int N>
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  constexpr uint32_t COLS = N <= 64 ? 64 : N <= 128 ? 128 : 256;
  constexpr uint32_t IDESC_N = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  const int tid = threadIdx.x;
  // pseudo-random operand bytes: the power a tensor-core loop draws (and with it the sustained clock) depends on the data
  for (int e = tid; e < (TM + N) * 128 / 16; e += 128) {
    uint32_t h = (uint32_t)e * 2654435761u + 0x9e3779b9u * (blockIdx.x + 1);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      h ^= h << 13;
      h ^= h >> 17;
      h ^= h << 5;
      w[q] = h;
    }
    reinterpret_cast<uint4*>(smem)[e] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  if (tid == 0) {
    auto desc = [](uint32_t addr) {
      return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    };
    const uint64_t da = desc(smem_u32(smem)), db = desc(smem_u32(smem + TM * 128));
    for (int i = 0; i < iters; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                     "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(IDESC_N), "r"((i | k) ? 1u : 0u)
                     : "memory");
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  tc_fence_before();
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS));
}

}  // namespace oz
}  // namespace nls
