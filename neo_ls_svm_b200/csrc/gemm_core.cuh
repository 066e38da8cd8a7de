// gemm_core.cuh — the FP64 tensor-core mainloop shared by every GEMM-shaped stage of the hot path.
//
// All stages are brought to the same "TN" form  C[M,N] = sum_k A[M,k] * B[N,k]  with BOTH operands
// K-contiguous in global memory, so one mainloop serves the feature map (X.W), the Hermitian Gram
// (phi^H S^2 phi), the eigenbasis projection (phi.Q), the gamma sweep ([P|U].r) and predict_std.
//
//  * CTA tile 128 (M) x 64 (N), K step 16 doubles (= one 128-byte swizzle row) per pipeline stage.
//  * 8 consumer warps (4 along M x 2 along N), each owning a 32x32 warp tile = 4x4 DMMA.8x8x4 tiles,
//    with up to two accumulator sets (real/imag or numerator/denominator) held in registers —
//    tcgen05/TMEM has no FP64 kind, so DMMA + register accumulators is the sm_100a FP64 tensor path.
//  * One elected thread streams operand tiles with TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into
//    a 4-stage ring guarded by full/empty mbarriers, two k-tiles ahead of the math.
//  * Fragment loads are conflict-free: with the 128B swizzle and the k assignment
//    {2c, 2c+1, 2c+8, 2c+9} per MMA step, each half-warp of an 8x4 (or 4x8) fragment load hits 16
//    distinct eight-byte bank pairs (the 2-wavefront minimum for a 256-byte warp load).
//
// Three arithmetic modes (operand "planes" are separate TMA boxes):
//   MODE_REAL     R += xa.ya                                   (1 A plane, 1 B plane)
//   MODE_DUAL_A   R += xa.ya ;  I += xb.ya                     (2 A planes, 1 B plane)
//   MODE_COMPLEX  R += xa.ya + xb.yb ;  I += xb.ya - xa.yb     (2 A planes, 2 B planes)
//   MODE_PAIR     R += xa.ya ;  I += xb.yb                     (2 A planes, 2 B planes; two GEMMs at once)
// MODE_COMPLEX is conj(xa + i xb) * (ya + i yb) up to the sign of I, i.e. one complex GEMM on a real
// MMA with each operand tile loaded once.
#pragma once
#include "ptx.cuh"

namespace nls {

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 16;
constexpr int STAGES = 4;
constexpr int PREFETCH = 2;  // k-tiles in flight ahead of the math (STAGES - PREFETCH - 1 = slack)
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
constexpr int GEMM_THREADS = CONSUMER_THREADS;
constexpr int A_PLANE_BYTES = BM * BK * 8;  // 16 KiB
constexpr int B_PLANE_BYTES = BN * BK * 8;  //  8 KiB

enum { MODE_REAL = 0, MODE_DUAL_A = 1, MODE_COMPLEX = 2, MODE_PAIR = 3 };

template <int MODE>
struct ModeTraits {
  static constexpr int A_PLANES = (MODE == MODE_REAL) ? 1 : 2;
  static constexpr int B_PLANES = (MODE == MODE_COMPLEX || MODE == MODE_PAIR) ? 2 : 1;
  static constexpr int STAGE_BYTES = A_PLANES * A_PLANE_BYTES + B_PLANES * B_PLANE_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  // 1 KiB alignment slack + pipeline + barriers.
  static constexpr int SMEM_BYTES = 1024 + PIPE_BYTES + 256;
};

// Raw operand description (used by the non-TMA reference loader and to build tensor maps).
struct Operand {
  const double* ptr;  // element (row, k) at ptr[row * ld + k]
  long long ld;       // row pitch in doubles (even)
  int rows;           // valid rows (others read as zero)
  int kext;           // valid K extent (others read as zero)
  int plane_drow;     // row offset of plane b relative to plane a
  int plane_dk;       // K offset of plane b relative to plane a
};

struct Tile {
  int m0, n0;           // first A row / first B row of this CTA's tile
  int k_begin, k_end;   // K range, k_begin multiple of BK
  bool valid;
};

struct Acc {
  double r[4][4][2];
  double i[4][4][2];
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Swizzled byte offset of element (row, k) in a [rows][16] plane (CU_TENSOR_MAP_SWIZZLE_128B).
__device__ __forceinline__ uint32_t swz(int row, int k) {
  return row * 128 + ((((k >> 1) ^ (row & 7)) & 7) << 4) + ((k & 1) << 3);
}

template <int MODE>
__device__ __forceinline__ void mma_stage(Acc& acc, uint32_t stage_base, int warp_m, int warp_n, int lane) {
  using T = ModeTraits<MODE>;
  const uint32_t a0 = stage_base + (warp_m * 32) * 128;
  const uint32_t a1 = a0 + A_PLANE_BYTES;
  const uint32_t b0 = stage_base + T::A_PLANES * A_PLANE_BYTES + (warp_n * 32) * 128;
  const uint32_t b1 = b0 + B_PLANE_BYTES;
  const int r8 = lane >> 2;
  const int q = lane & 3;
#pragma unroll
  for (int kk = 0; kk < BK / 4; ++kk) {
    // MMA step kk contracts k = {2kk, 2kk+1, 2kk+8, 2kk+9} of the stage (any order of the reduction
    // index is valid as long as A and B agree).  The two k-pairs of a lane quad then sit in 16-byte
    // chunks kk and kk+4, so the 16 lanes of each half-warp (rows r8..r8+3, XOR-swizzled) hit 16
    // distinct 8-byte bank pairs: one conflict-free wavefront per half-warp.
    const uint32_t lo = r8 * 128 + ((((kk + 4 * (q >> 1)) ^ r8) & 7) << 4) + ((q & 1) << 3);
    double xa[4], xb[4], ya[4], yb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      xa[t] = lds_f64(a0 + t * 1024 + lo);
      ya[t] = lds_f64(b0 + t * 1024 + lo);
      if (T::A_PLANES == 2) xb[t] = lds_f64(a1 + t * 1024 + lo);
      if (T::B_PLANES == 2) yb[t] = lds_f64(b1 + t * 1024 + lo);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(acc.r[i][j][0], acc.r[i][j][1], xa[i], ya[j]);
    if (MODE == MODE_DUAL_A) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc.i[i][j][0], acc.i[i][j][1], xb[i], ya[j]);
    }
    if (MODE == MODE_PAIR) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc.i[i][j][0], acc.i[i][j][1], xb[i], yb[j]);
    }
    if (MODE == MODE_COMPLEX) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc.r[i][j][0], acc.r[i][j][1], xb[i], yb[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc.i[i][j][0], acc.i[i][j][1], xb[i], ya[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i) xa[i] = -xa[i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc.i[i][j][0], acc.i[i][j][1], xa[i], yb[j]);
    }
  }
}

// Reference loader (no TMA): consumer threads copy one stage with bounds checks into the same
// swizzled layout.  Slow; exists so that the TMA pipeline can be validated on-device against it.
template <int MODE>
__device__ __forceinline__ void load_stage_plain(uint8_t* stage, const Operand& A, const Operand& B, const Tile& t,
                                                 int k0, int tid) {
  using T = ModeTraits<MODE>;
#pragma unroll 1
  for (int pl = 0; pl < T::A_PLANES; ++pl) {
    uint8_t* dst = stage + pl * A_PLANE_BYTES;
    for (int e = tid; e < BM * BK; e += CONSUMER_THREADS) {
      int row = e / BK, k = e % BK;
      long long gr = (long long)t.m0 + row + (pl ? A.plane_drow : 0);
      long long gk = (long long)k0 + k + (pl ? A.plane_dk : 0);
      double v = 0.0;
      if (gr >= 0 && gr < A.rows && k0 + k < A.kext) v = A.ptr[gr * A.ld + gk];
      *reinterpret_cast<double*>(dst + swz(row, k)) = v;
    }
  }
#pragma unroll 1
  for (int pl = 0; pl < T::B_PLANES; ++pl) {
    uint8_t* dst = stage + T::A_PLANES * A_PLANE_BYTES + pl * B_PLANE_BYTES;
    for (int e = tid; e < BN * BK; e += CONSUMER_THREADS) {
      int row = e / BK, k = e % BK;
      long long gr = (long long)t.n0 + row + (pl ? B.plane_drow : 0);
      long long gk = (long long)k0 + k + (pl ? B.plane_dk : 0);
      double v = 0.0;
      if (gr >= 0 && gr < B.rows && k0 + k < B.kext) v = B.ptr[gr * B.ld + gk];
      *reinterpret_cast<double*>(dst + swz(row, k)) = v;
    }
  }
}

// Op must provide:
//   struct Params { Operand A, B; ... };
//   static __device__ Tile tile(const Params&);                       (from blockIdx)
//   static __device__ void epilogue(const Params&, const Tile&, Acc&, int warp_m, int warp_n,
//                                   int lane, uint8_t* scratch);
//
// Pipeline: thread 0 doubles as the TMA producer.  Before consuming k-tile kt it refills the slot
// that k-tile kt + PREFETCH will use, which every warp released PREFETCH - STAGES iterations ago
// (one iteration of slack), so the refill never stalls the math unless a warp lags a full stage.
// No dedicated producer warp: 256 threads leave the full 255-register budget to the accumulators.
template <int MODE, class Op, bool USE_TMA>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            const typename Op::Params p) {
  using T = ModeTraits<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T::PIPE_BYTES);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = full0 + STAGES * 8;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const Tile t = Op::tile(p);
  if (!t.valid) return;
  const int num_kt = (t.k_end - t.k_begin + BK - 1) / BK;

  auto issue = [&](int kt) {
    const int s = kt % STAGES;
    const uint32_t full = full0 + s * 8;
    mbar_arrive_expect_tx(full, T::STAGE_BYTES);
    const uint32_t dst = smem_base + s * T::STAGE_BYTES;
    const int k = t.k_begin + kt * BK;
    tma_load_2d(dst, &mapA, k, t.m0, full);
    if (T::A_PLANES == 2) tma_load_2d(dst + A_PLANE_BYTES, &mapA, k + p.A.plane_dk, t.m0 + p.A.plane_drow, full);
    const uint32_t dstb = dst + T::A_PLANES * A_PLANE_BYTES;
    tma_load_2d(dstb, &mapB, k, t.n0, full);
    if (T::B_PLANES == 2) tma_load_2d(dstb + B_PLANE_BYTES, &mapB, k + p.B.plane_dk, t.n0 + p.B.plane_drow, full);
  };

  if (USE_TMA) {
    if (tid == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapB);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full0 + s * 8, 1);
        mbar_init(empty0 + s * 8, CONSUMER_WARPS);
      }
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
      for (int kt = 0; kt < PREFETCH && kt < num_kt; ++kt) issue(kt);
    }
  }

  const int warp_m = warp & 3;
  const int warp_n = warp >> 2;
  Acc acc;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc.r[i][j][0] = acc.r[i][j][1] = 0.0;
      acc.i[i][j][0] = acc.i[i][j][1] = 0.0;
    }

  if (USE_TMA) {
#pragma unroll 1
    for (int kt = 0; kt < num_kt; ++kt) {
      if (tid == 0 && kt + PREFETCH < num_kt) {
        const int nk = kt + PREFETCH;
        // Slot nk % STAGES was last filled with k-tile nk - STAGES; wait until all warps released it.
        if (nk >= STAGES) mbar_wait(empty0 + (nk % STAGES) * 8, ((nk / STAGES) - 1) & 1);
        issue(nk);
      }
      __syncwarp();
      const int s = kt % STAGES;
      mbar_wait(full0 + s * 8, (kt / STAGES) & 1);
      mma_stage<MODE>(acc, smem_base + s * T::STAGE_BYTES, warp_m, warp_n, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + s * 8);
    }
  } else {
#pragma unroll 1
    for (int kt = 0; kt < num_kt; ++kt) {
      __syncthreads();
      load_stage_plain<MODE>(smem, p.A, p.B, t, t.k_begin + kt * BK, tid);
      __syncthreads();
      mma_stage<MODE>(acc, smem_base, warp_m, warp_n, lane);
    }
  }
  // Every TMA write has been consumed; the pipeline memory is reusable as epilogue scratch once all
  // warps are past the mainloop.
  __syncthreads();
  Op::epilogue(p, t, acc, warp_m, warp_n, lane, smem);
}

}  // namespace nls
