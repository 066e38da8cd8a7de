// Micro-probe for DESIGN.md §9 (1): one CTA issues tcgen05.mma kind::i8 (INT8 x INT8 -> INT32 in TMEM) on operands that
// its threads wrote to shared memory in the canonical K-major SWIZZLE_128B layout, reads the accumulator back with
// tcgen05.ld and checks it against the host; then times a train of MMAs to get the per-SM INT8 issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neo_ls_svm_b200/csrc -o scripts/micro/i8_umma.bin scripts/micro/i8_umma.cu
// Not product code: it establishes that the digit-plane products of scripts/ozaki_study.py map onto the sm_100a tensor
// cores, and what one SM sustains.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "ptx.cuh"
using namespace nls;

constexpr int M = 128, KB = 128, UMMA_K = 32;  // tile rows, bytes (= int8 elements) per K slab, K per MMA

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units            [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused: swizzled)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows x 128 B       [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)           [46,48)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B                             [61,64)
  return d;
}

template <int N, bool B_UNSIGNED>
__global__ void __launch_bounds__(128) i8_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                 int32_t* __restrict__ C, int K, int reps, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;             // M x 128 B
  uint8_t* sB = smem + M * KB;    // N x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : N <= 256 ? 256 : 512;  // power of two >= 32
  // instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N, M
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | ((B_UNSIGNED ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_holder;
  uint32_t phase = 0;

  auto issue = [&](bool accumulate_first) {
#pragma unroll
    for (int k = 0; k < KB / UMMA_K; ++k) {
      const uint64_t da = umma_desc_k_sw128(smem_u32(sA)) + (uint64_t)((k * UMMA_K) >> 4);
      const uint64_t db = umma_desc_k_sw128(smem_u32(sB)) + (uint64_t)((k * UMMA_K) >> 4);
      const uint32_t acc = (accumulate_first || k > 0) ? 1u : 0u;
      asm volatile(
          "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(IDESC), "r"(acc)
          : "memory");
    }
  };
  auto commit = [&]() {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  };

  for (int slab = 0; slab < K / KB; ++slab) {
    // Canonical K-major SWIZZLE_128B tile: row r at r * 128 B, its 16-byte chunk c at ((c ^ (r & 7)) * 16).
    for (int e = tid; e < M * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + slab * KB + c * 16);
      *reinterpret_cast<uint4*>(sA + r * KB + ((c ^ (r & 7)) << 4)) = v;
    }
    for (int e = tid; e < N * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * K + slab * KB + c * 16);
      *reinterpret_cast<uint4*>(sB + r * KB + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      issue(slab > 0);
      commit();
    }
    mbar_wait(smem_u32(&bar), phase);  // the MMAs have read this slab: shared memory may be overwritten
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // Accumulator: row i of D lives in TMEM lane i, column j in TMEM column j; warp w may touch lanes 32 w .. 32 w + 31.
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) C[(size_t)(32 * warp + lane) * N + c0 + j] = (int32_t)r[j];
  }

  // Issue-rate probe: `reps` trains of 4 MMAs on the resident tiles, one commit at the end.
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (reps > 0) {
    long long t0 = 0;
    if (tid == 0) {
      t0 = clock64();
      for (int i = 0; i < reps; ++i) issue(true);
      commit();
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    if (tid == 0) cyc[0] = clock64() - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS));
}

template <int N, bool B_UNSIGNED>
static int run(int K, int reps) {
  std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
  srand(7 + N);
  for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
  int8_t *dA, *dB;
  int32_t* dC;
  long long* dcyc;
  cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dC, (size_t)M * N * 4); cudaMalloc(&dcyc, 8);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  cudaMemset(dC, 0xff, (size_t)M * N * 4);
  const size_t smem = 1024 + (size_t)(M + N) * KB;
  cudaFuncSetAttribute(i8_kernel<N, B_UNSIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  i8_kernel<N, B_UNSIGNED><<<1, 128, smem>>>(dA, dB, dC, K, reps, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d: CUDA error: %s\n", N, cudaGetErrorString(e));
    return 1;
  }
  std::vector<int32_t> C((size_t)M * N);
  long long cyc = 0;
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
  long long bad = 0, first = -1;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k)
        ref += (int32_t)A[(size_t)i * K + k] * (B_UNSIGNED ? (int32_t)(uint8_t)B[(size_t)j * K + k] : (int32_t)B[(size_t)j * K + k]);
      if (ref != C[(size_t)i * N + j]) {
        if (first < 0) first = (long long)i * N + j;
        ++bad;
      }
    }
  printf("N=%d K=%d B %s: %lld of %d entries differ from the host INT32 product", N, K, B_UNSIGNED ? "unsigned" : "signed", bad, M * N);
  if (bad) printf(" (first at row %lld col %lld: got %d)", first / N, first % N, C[first]);
  const double ops = 2.0 * M * N * KB * (double)reps;
  printf("; %d x 4 MMAs in %lld cycles = %.0f INT8 op/cycle/SM\n", reps, cyc, cyc > 0 ? ops / (double)cyc : 0.0);
  return bad != 0;
}

int main() {
  int rc = 0;
  rc |= run<64, false>(256, 2000);
  rc |= run<256, false>(256, 2000);
  rc |= run<80, true>(256, 2000);  // radix-256 planes: signed x unsigned digits, N = 80 (6 accumulators fit TMEM)
  return rc;
}
