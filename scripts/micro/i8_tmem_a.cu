// Probe: tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY (tcgen05.cp smem -> TMEM), for the Ozaki tile.
// At M = 128, N = 64 with both operands in shared memory every MMA reads 6 KB (48 cycles at 128 B/clk) for 32 cycles of
// math.  With A in TMEM only B (N x 32 B) comes from shared memory.  Checks, for both tile images of csrc/ozaki.cuh:
//   1. D = A B^T with A from shared memory (reference) and with A copied to TMEM by tcgen05.cp.128x256b, against the host;
//   2. cycles per "k step" of 28 MMAs (7 A planes x 7 B planes, p + q <= 6) at N = 48 and N = 64, A from smem / from TMEM,
//      with and without the 7 tcgen05.cp of the next step's A planes in the stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I neo_ls_svm_b200/csrc -o scripts/micro/i8_tmem_a.bin scripts/micro/i8_tmem_a.cu
// Not product code.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ozaki.cuh"
using namespace nls;

constexpr int M = 128, KSTEP = 32, S = 7;

template <int LAYOUT>
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint32_t smem_addr) {
  const uint64_t d = oz::smem_desc<LAYOUT>(smem_addr);
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(d) : "memory");
}
template <int N>
__device__ __forceinline__ void mma_a_tmem(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t acc) {
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(IDESC), "r"(acc)
               : "memory");
}
template <int N>
__device__ __forceinline__ void mma_a_smem(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t acc) {
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(acc)
               : "memory");
}

// Warp-convergent issue: all 32 lanes run the (uniform) address arithmetic, one elected lane issues the instruction.
template <int N>
__device__ __forceinline__ void mma_a_tmem_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc) {
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  asm volatile("{\n .reg .pred p, q;\n elect.sync _|q, 0xffffffff;\n setp.ne.b32 p, 1, 0;\n @q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(IDESC)
               : "memory");
}
template <int N>
__device__ __forceinline__ void mma_a_smem_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc) {
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  asm volatile("{\n .reg .pred p, q;\n elect.sync _|q, 0xffffffff;\n setp.ne.b32 p, 1, 0;\n @q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(IDESC)
               : "memory");
}

// out[mode][row][col]: mode 0 = A from smem, mode 1 = A from TMEM.  cyc[0..3]: smem-A, TMEM-A, TMEM-A + cp, cp only.
template <int LAYOUT, int N>
__global__ void __launch_bounds__(128) probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ out,
                                                    int reps, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                      // S planes x 128 x 32 B
  uint8_t* sB = smem + S * M * KSTEP;      // S planes x N x 32 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t ACC_COLS = S * N;     // 7 accumulators
  constexpr uint32_t A_COL0 = ACC_COLS;    // then two A buffers of S * 8 columns
  constexpr uint32_t A_BUF = S * 8;
  constexpr bool DOUBLE = ACC_COLS + 2 * A_BUF <= 512;  // N = 64 has room for one A buffer only
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  // plane p of both operands = the same matrix shifted by p (contents only matter for plane 0 in the correctness check)
  for (int e = tid; e < S * M * 2; e += 128) {
    const int p = e / (M * 2), r = (e >> 1) % M, c = e & 1;
    uint4 v = *reinterpret_cast<const uint4*>(A + r * KSTEP + c * 16);
    *reinterpret_cast<uint4*>(sA + p * M * KSTEP + oz::tile_off<LAYOUT>(r, c)) = v;
  }
  for (int e = tid; e < S * N * 2; e += 128) {
    const int p = e / (N * 2), r = (e >> 1) % N, c = e & 1;
    uint4 v = *reinterpret_cast<const uint4*>(B + r * KSTEP + c * 16);
    *reinterpret_cast<uint4*>(sB + p * N * KSTEP + oz::tile_off<LAYOUT>(r, c)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  oz::tc_fence_before();
  __syncthreads();
  oz::tc_fence_after();
  const uint32_t tmem = tmem_holder;
  uint32_t phase = 0;
  const uint64_t da = oz::smem_desc<LAYOUT>(smem_u32(sA)), db = oz::smem_desc<LAYOUT>(smem_u32(sB));
  for (int mode = 0; mode < 2; ++mode) {
    if (tid == 0) {
      if (mode == 0) {
        mma_a_smem<N>(tmem, da, db, 0u);
      } else {
        tmem_cp_128x256b<LAYOUT>(tmem + A_COL0, smem_u32(sA));
        mma_a_tmem<N>(tmem, tmem + A_COL0, db, 0u);
      }
      oz::umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    oz::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      oz::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
      oz::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (blockIdx.x == 0) out[(mode * M + 32 * warp + lane) * N + c0 + j] = (int32_t)r[j];
    }
    oz::tc_fence_before();
    __syncthreads();
    oz::tc_fence_after();
  }
  // timing
  for (int variant = 0; variant < 4; ++variant) {
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t abuf = tmem + A_COL0 + (uint32_t)((DOUBLE ? (i & 1) : 0) * A_BUF);
        if (variant >= 2)
          for (int p = 0; p < S; ++p) tmem_cp_128x256b<LAYOUT>(tmem + A_COL0 + (uint32_t)((DOUBLE ? ((i + 1) & 1) : 0) * A_BUF) + p * 8, smem_u32(sA + p * M * KSTEP));
        if (variant < 3) {
#pragma unroll
          for (int p = 0; p < S; ++p)
#pragma unroll
            for (int q = 0; p + q < S; ++q) {
              if (variant == 0)
                mma_a_smem<N>(tmem + (uint32_t)((p + q) * N), da + (uint64_t)((p * M * KSTEP) >> 4), db + (uint64_t)((q * N * KSTEP) >> 4), 1u);
              else
                mma_a_tmem<N>(tmem + (uint32_t)((p + q) * N), abuf + p * 8, db + (uint64_t)((q * N * KSTEP) >> 4), 1u);
            }
        }
      }
      oz::umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    if (tid == 0 && blockIdx.x == 0) cyc[variant] = clock64() - t0;
    oz::tc_fence_after();
    __syncthreads();
  }
  // variants 4 / 5: the same 28-MMA steps issued warp-convergently (smem-A / TMEM-A)
  for (int variant = 4; variant < 6; ++variant) {
    long long t0 = 0;
    if (warp == 0) {
      t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint32_t abuf = tmem + A_COL0 + (uint32_t)((DOUBLE ? (i & 1) : 0) * A_BUF);
#pragma unroll
        for (int p = 0; p < S; ++p)
#pragma unroll
          for (int q = 0; p + q < S; ++q) {
            if (variant == 4)
              mma_a_smem_w<N>(tmem + (uint32_t)((p + q) * N), da + (uint64_t)((p * M * KSTEP) >> 4), db + (uint64_t)((q * N * KSTEP) >> 4));
            else
              mma_a_tmem_w<N>(tmem + (uint32_t)((p + q) * N), abuf + p * 8, db + (uint64_t)((q * N * KSTEP) >> 4));
          }
      }
      __syncwarp();
      if (lane == 0) oz::umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    if (tid == 0 && blockIdx.x == 0) cyc[variant] = clock64() - t0;
    oz::tc_fence_after();
    __syncthreads();
  }
  oz::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int LAYOUT, int N>
static int run() {
  std::vector<int8_t> A(M * KSTEP), B(N * KSTEP);
  srand(3 + N);
  for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB;
  int32_t* dO;
  long long* dC;
  cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dO, 2 * M * N * 4); cudaMalloc(&dC, 6 * 8);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, 2 * M * N * 4);
  const int smem = 1024 + S * (M + N) * KSTEP;
  cudaFuncSetAttribute(probe_kernel<LAYOUT, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2000;
  probe_kernel<LAYOUT, N><<<148, 128, smem>>>(dA, dB, dO, reps, dC);  // a full grid: single-CTA runs are issue-throttled
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("layout %d N %d: CUDA error %s\n", LAYOUT, N, cudaGetErrorString(e));
    return 1;
  }
  std::vector<int32_t> O(2 * M * N);
  long long cyc[6];
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc, dC, sizeof cyc, cudaMemcpyDeviceToHost);
  int bad[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode)
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        int ref = 0;
        for (int k = 0; k < KSTEP; ++k) ref += (int)A[i * KSTEP + k] * (int)B[j * KSTEP + k];
        if (O[(mode * M + i) * N + j] != ref) ++bad[mode];
      }
  printf("layout %d N %2d: wrong entries: A from smem %d, A from TMEM %d of %d;  cycles per 28-MMA k step: smem-A %.0f, TMEM-A %.0f, TMEM-A + 7 cp %.0f, 7 cp alone %.0f; warp-convergent issue: smem-A %.0f, TMEM-A %.0f\n",
         LAYOUT, N, bad[0], bad[1], M * N, cyc[0] / (double)reps, cyc[1] / (double)reps, cyc[2] / (double)reps, cyc[3] / (double)reps, cyc[4] / (double)reps, cyc[5] / (double)reps);
  return bad[0] || bad[1];
}

int main() {
  int rc = 0;
  rc |= run<0, 48>();
  rc |= run<6, 48>();
  rc |= run<0, 64>();
  rc |= run<6, 64>();
  printf("rc = %d\n", rc);
  return 0;
}
