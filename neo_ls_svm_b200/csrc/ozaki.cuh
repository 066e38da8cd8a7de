// ozaki.cuh — FP64-accurate GEMM on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators) by the Ozaki scheme.
//
// sm_100a has no FP64 kind on tcgen05; DMMA tops out at 128 flop/cycle/SM.  The eigenbasis projection T = phi Q
// (reference: _neo_ls_svm.py:134, :137 — 44 % of the C3 fit) has bounded operands (|phi| <= D^-1/2, Q unitary), so
// each operand row is scaled by a power of two to |x| <= 1/2, rounded ONCE to 49 fractional bits and cut exactly
// into S = 7 signed base-128 digits x = 2^e sum_p d_p 128^-(p+1).  A digit-plane product A_p B_q^T is an exact
// INT8 x INT8 -> INT32 GEMM; the S(S+1)/2 = 28 products with p + q <= 6 go into 7 TMEM accumulators (one per
// level t = p + q, 7 x 64 = 448 of the 512 columns) and are recombined in FP64 in the epilogue, smallest level
// first.  Dropped levels (p + q >= 7) are below 2^-47 of max|a| max|b| per term.
//
// Data path (everything is laid out by our own slicing kernels, so no tensor maps are needed):
//   * planes live in global memory TILE-MAJOR: [row block][k step][plane][rows x 32 bytes], each [rows x 32 B] tile
//     already in the canonical K-major SWIZZLE_32B shared-memory image.  A pipeline stage (one k step of 32: 7 A
//     tiles of 128 rows + 7 B tiles of 64 rows = 42 KB) is two contiguous blocks, fetched with two 1-D
//     cp.async.bulk copies that complete on the stage's mbarrier.
//   * persistent CTAs (one per SM), warp-specialised: warp 0 = bulk-copy producer (one lane), warp 1 = MMA issuer
//     (one lane, 28 tcgen05.mma per stage, tcgen05.commit releases the stage), warps 2-9 = epilogue (tcgen05.ld,
//     FP64 recombination, fused epilogue functor).  5 stages in flight; the accumulators are handed back to the
//     MMA warp as soon as they are drained into registers, so the epilogue math and stores of tile i overlap the
//     mainloop of tile i + 1.
//   * B tile rows 0..31 are the "real" rows and rows 32..63 the "imaginary" rows of the same 32 complex columns,
//     so one thread owns Re and Im of its (row, column) pairs.
#pragma once
#include "ptx.cuh"

namespace nls {
namespace oz {

constexpr int S = 7;               // digit planes per operand (7 bits each)
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
constexpr int THREADS = 320;       // producer warp, MMA warp, 8 epilogue warps
constexpr int FRAC_BITS = 7 * S;   // 49
constexpr uint32_t TMEM_COLS = 512;
constexpr int IMAGE = 6;           // tile image the product uses: canonical K-major SWIZZLE_32B (see tile_off)
constexpr int MAX_K = 65536;       // INT32 accumulators: (t + 1) 64^2 K < 2^31 for every level t <= 6

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
// x 2^-e in [-1/2, 1/2]  ->  V = rint(x 2^(49 - e))  =  sum_p d_p 128^(6 - p),  d_p in [-64, 63] (d_0 up to 64).
// `scale` = 2^(49 - e).  The rounding happens once, in the FMA against 1.5 2^52 (round to nearest even).
__host__ __device__ __forceinline__ long long quantise(double x, double scale) {
#ifdef __CUDA_ARCH__
  const double t = fma(x, scale, 6755399441055744.0);
  return __double_as_longlong(t) - 0x4338000000000000LL;
#else
  return (long long)__builtin_rint(x * scale);
#endif
}
constexpr long long DIGIT_BIAS = 64LL * (1 + 128LL + 128LL * 128 + 128LL * 128 * 128 + 128LL * 128 * 128 * 128 +
                                         128LL * 128 * 128 * 128 * 128 + 128LL * 128 * 128 * 128 * 128 * 128);
__host__ __device__ __forceinline__ int digit(long long v_biased, int p) {  // p = 0 is the leading digit
  const long long u = v_biased >> (7 * (S - 1 - p));
  return (int)(p == 0 ? u : (u & 127)) - 64;
}

// Power-of-two exponent e with |x| 2^-e <= 1/2 for all |x| <= amax (0 when amax == 0).
__host__ __device__ __forceinline__ int scale_exponent(double amax) {
  if (!(amax > 0.0)) return 0;
  int ex;
  const double f = frexp(amax, &ex);  // amax = f 2^ex, f in [1/2, 1)
  return f == 0.5 ? ex : ex + 1;
}

// ---- slicing kernels ------------------------------------------------------------------------------
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
    const int k0 = ks * KS + c * 16;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[p][j] = 0;
    if (row < rows) {
      const double* src = X + row * ld + k0;
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const int col = (k0 + j) % Dp;  // Dp is a multiple of 16: a chunk never straddles the two halves
        double2 v = make_double2(0.0, 0.0);
        if (col < D) v = *reinterpret_cast<const double2*>(src + j);
        if (col + 1 >= D) v.y = 0.0;
        const long long q0 = quantise(v.x, scale) + DIGIT_BIAS, q1 = quantise(v.y, scale) + DIGIT_BIAS;
#pragma unroll
        for (int p = 0; p < S; ++p) {
          const uint32_t b0 = (uint32_t)(digit(q0, p) & 0xff), b1 = (uint32_t)(digit(q1, p) & 0xff);
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
// becomes two rows of the real product over k in [0, 2 Dp):   re row = [ ya_j | yb_j ],  im row = [ -yb_j | ya_j ]
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
    const int k0 = ks * KS + c * 16;
    const int half = k0 / Dp, col0 = k0 % Dp;
    uint32_t w[S][4];
#pragma unroll
    for (int p = 0; p < S; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) w[p][q] = 0;
    if (j < cols) {
      // re row: half 0 -> ya, half 1 -> yb;   im row: half 0 -> -yb, half 1 -> ya
      const bool use_b = (half == 1) != im;
      const double sign = (im && half == 0) ? -1.0 : 1.0;
      const double* src = Bt + (long long)((use_b ? Np : 0) + j) * Dp + col0;
      const double scale = sign * ldexp(1.0, FRAC_BITS - ex[j]);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const double v = col0 + i < D ? src[i] : 0.0;
        const long long q = quantise(v, scale) + DIGIT_BIAS;
#pragma unroll
        for (int p = 0; p < S; ++p) w[p][i >> 2] |= (uint32_t)(digit(q, p) & 0xff) << (8 * (i & 3));
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
};

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
  static __device__ __forceinline__ void apply(const Params& p, long long row, int col0, const double (&sr)[16], const double (&si)[16]) {
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

struct EpiStore {  // raw T planes (probe / tests): Tr[row][col] = sr * colscale, Ti likewise
  struct Params {
    int n_rows, m;
    const double* colscale;
    double* Tr;
    double* Ti;
    long long ld;
  };
  static __device__ __forceinline__ void apply(const Params& p, long long row, int col0, const double (&sr)[16], const double (&si)[16]) {
    if (row >= p.n_rows) return;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < p.m) {
        p.Tr[row * p.ld + col0 + j] = sr[j] * p.colscale[col0 + j];
        p.Ti[row * p.ld + col0 + j] = si[j] * p.colscale[col0 + j];
      }
  }
};

template <int LAYOUT, class Epi>
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
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    mbar_init(smem_u32(&tmem_empty_bar), 8);  // one arrival per epilogue warp
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint32_t smem_base = smem_u32(smem);
  const long long total_tiles = (long long)g.row_blocks * g.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long rb = tile / g.n_tiles;
        const int nb = (int)(tile % g.n_tiles);
        const int8_t* a = g.A + rb * g.nks * (long long)A_STAGE;
        const int8_t* b = g.B + (long long)nb * g.nks * (long long)B_STAGE;
        for (int ks = 0; ks < g.nks; ++ks, ++it) {
          const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_arrive_expect_tx(fb, STAGE_BYTES);
          bulk_g2s(smem_base + s * STAGE_BYTES, a + (long long)ks * A_STAGE, A_STAGE, fb);
          bulk_g2s(smem_base + s * STAGE_BYTES + A_STAGE, b + (long long)ks * B_STAGE, B_STAGE, fb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tile_i = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
        mbar_wait(smem_u32(&tmem_empty_bar), (tile_i & 1) ^ 1);  // accumulators drained by the epilogue warps
        tc_fence_after();
        for (int ks = 0; ks < g.nks; ++ks, ++it) {
          const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint64_t da = smem_desc<LAYOUT>(smem_base + s * STAGE_BYTES);
          const uint64_t db = smem_desc<LAYOUT>(smem_base + s * STAGE_BYTES + A_STAGE);
          const uint32_t acc0 = ks > 0 ? 1u : 0u;
#pragma unroll
          for (int p = 0; p < S; ++p)
#pragma unroll
            for (int q = 0; p + q < S; ++q) {
              // level t = p + q; its first product of the tile (p = 0) overwrites, everything else accumulates
              umma_i8(tmem + (uint32_t)((p + q) * TN), da + (uint64_t)((p * A_TILE) >> 4), db + (uint64_t)((q * B_TILE) >> 4),
                      p == 0 ? acc0 : 1u);
            }
          umma_commit(smem_u32(&empty_bar[s]));  // stage reusable once these MMAs have read it
        }
        umma_commit(smem_u32(&tmem_full_bar));
      }
    }
  } else {
    const int e = warp - 2;
    const int quad = warp & 3;       // the TMEM lane quadrant this warp may access
    const int cb = (e >> 2) * 16;    // first of its 16 complex columns within the tile
    uint32_t tile_i = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      const long long rb = tile / g.n_tiles;
      const int nb = (int)(tile % g.n_tiles);
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
        const double w = 1.0 / (double)(1ull << (7 * (t + 2)));  // 128^-(t+2)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          sr[j] = fma(i2d(ar[j]), w, sr[j]);
          si[j] = fma(i2d(ai[j]), w, si[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar));
      Epi::apply(ep, rb * TM + quad * 32 + lane, nb * (TN / 2) + cb, sr, si);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace oz
}  // namespace nls
