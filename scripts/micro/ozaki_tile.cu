// Micro-probe for DESIGN.md §9 (1), second step: one FP64 tile product C = A B^T (128 x 64 x K) computed on the INT8
// tensor cores by the Ozaki scheme — operands cut on the fly into 7 signed 7-bit digit planes in shared memory (K-major
// SWIZZLE_128B), 28 plane products per K slab issued as tcgen05.mma kind::i8 into 7 TMEM accumulators (one per p + q),
// recombined in FP64 in the epilogue — and compared with a long-double host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neo_ls_svm_b200/csrc -o scripts/micro/ozaki_tile.bin scripts/micro/ozaki_tile.cu
// Not product code (see scripts/ozaki_study.py for the parity study this mirrors).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include "ptx.cuh"
using namespace nls;

constexpr int M = 128, N = 64, KB = 128, UMMA_K = 32, S = 7;  // S digit planes of 7 bits

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Digit planes of `rows` x 128 doubles (row r scaled by 2^-ex[r] to |x| <= 1/2): plane p, byte (r, k) at the
// SWIZZLE_128B position of a K-major tile.
__device__ __forceinline__ void slice_slab(const double* __restrict__ X, int ldx, int rows, const int* __restrict__ ex,
                                           uint8_t* __restrict__ planes, int plane_bytes, int tid) {
  for (int e = tid; e < rows * KB; e += 128) {
    const int r = e / KB, k = e % KB;
    double x = ldexp(X[(size_t)r * ldx + k], -ex[r]);
    const int off = r * KB + ((((k >> 4) ^ (r & 7)) << 4) | (k & 15));
#pragma unroll
    for (int p = 0; p < S; ++p) {
      x *= 128.0;
      const int q = __double2int_rn(x);
      x -= (double)q;
      planes[p * plane_bytes + off] = (uint8_t)(int8_t)q;
    }
  }
}

__global__ void __launch_bounds__(128) ozaki_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                    const int* __restrict__ exA, const int* __restrict__ exB,
                                                    double* __restrict__ C, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int PA = M * KB, PB = N * KB;  // bytes per digit plane
  uint8_t* sA = smem;                      // S planes of A
  uint8_t* sB = smem + S * PA;             // S planes of B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t COLS = 512;  // S accumulators of N columns = 448
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
  for (int slab = 0; slab < K / KB; ++slab) {
    slice_slab(A + slab * KB, K, M, exA, sA, PA, tid);
    slice_slab(B + slab * KB, K, N, exB, sB, PB, tid);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      uint32_t written = slab > 0 ? 0x7fu : 0u;  // accumulators that already hold a partial sum
      for (int p = 0; p < S; ++p)
        for (int q = 0; p + q < S; ++q) {
          const int t = p + q;
          for (int k = 0; k < KB / UMMA_K; ++k) {
            const uint64_t da = umma_desc_k_sw128(smem_u32(sA + p * PA)) + (uint64_t)((k * UMMA_K) >> 4);
            const uint64_t db = umma_desc_k_sw128(smem_u32(sB + q * PB)) + (uint64_t)((k * UMMA_K) >> 4);
            const uint32_t acc = ((written >> t) & 1u) | (k > 0 ? 1u : 0u);
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                    tmem + (uint32_t)(t * N)),
                "l"(da), "l"(db), "r"(IDESC), "r"(acc)
                : "memory");
          }
          written |= 1u << t;
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // Epilogue: C[row, col] = 2^(exA[row] + exB[col]) * sum_t 128^-(t+2) * acc_t[row, col], smallest terms first.
  const int row = 32 * warp + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    double sum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = 0.0;
    for (int t = S - 1; t >= 0; --t) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(t * N + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const double w = ldexp(1.0, -7 * (t + 2));
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] = fma((double)(int32_t)r[j], w, sum[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) C[(size_t)row * N + c0 + j] = ldexp(sum[j], exA[row] + exB[c0 + j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS));
}

static void exponents(const std::vector<double>& X, int rows, int K, std::vector<int>& ex) {
  ex.resize(rows);
  for (int r = 0; r < rows; ++r) {
    double amax = 0.0;
    for (int k = 0; k < K; ++k) amax = fmax(amax, fabs(X[(size_t)r * K + k]));
    ex[r] = amax > 0.0 ? (int)ceil(log2(amax)) + 1 : 0;  // |x| 2^-ex <= 1/2
  }
}

int main() {
  const int K = 512;
  std::vector<double> A((size_t)M * K), B((size_t)N * K);
  srand(3);
  auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * pow(10.0, -6.0 * rand() / (double)RAND_MAX); };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd() * 37.0;
  std::vector<int> exA, exB;
  exponents(A, M, K, exA);
  exponents(B, N, K, exB);
  double *dA, *dB, *dC;
  int *dexA, *dexB;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, (size_t)M * N * 8);
  cudaMalloc(&dexA, M * 4); cudaMalloc(&dexB, N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dexA, exA.data(), M * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dexB, exB.data(), N * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + (size_t)S * (M + N) * KB;
  cudaFuncSetAttribute(ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ozaki_kernel<<<1, 128, smem>>>(dA, dB, dexA, dexB, dC, K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<double> C((size_t)M * N);
  cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
  // Host emulation of the same scheme (same digits, exact integer plane sums, same recombination order): the GPU
  // result must agree with it bit for bit.
  std::vector<int8_t> dA_((size_t)S * M * K), dB_((size_t)S * N * K);
  auto digits = [&](const std::vector<double>& X, int rows, const std::vector<int>& ex, std::vector<int8_t>& out) {
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < K; ++k) {
        double x = ldexp(X[(size_t)r * K + k], -ex[r]);
        for (int p = 0; p < S; ++p) {
          x *= 128.0;
          const int q = (int)nearbyint(x);
          x -= (double)q;
          out[((size_t)p * rows + r) * K + k] = (int8_t)q;
        }
      }
  };
  digits(A, M, exA, dA_);
  digits(B, N, exB, dB_);
  long long mismatches = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      long long acc[S] = {0};
      for (int p = 0; p < S; ++p)
        for (int q = 0; p + q < S; ++q) {
          long long d = 0;
          for (int k = 0; k < K; ++k) d += (long long)dA_[((size_t)p * M + i) * K + k] * (long long)dB_[((size_t)q * N + j) * K + k];
          acc[p + q] += d;
        }
      double sum = 0.0;
      for (int t = S - 1; t >= 0; --t) sum = fma((double)acc[t], ldexp(1.0, -7 * (t + 2)), sum);
      if (ldexp(sum, exA[i] + exB[j]) != C[(size_t)i * N + j]) ++mismatches;
    }
  printf("bitwise agreement with the host emulation of the scheme: %lld of %d entries differ\n", mismatches, M * N);
  long double worst_oz = 0, worst_f64 = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      long double ref = 0, bound = 0;
      double f64 = 0;
      for (int k = 0; k < K; ++k) {
        ref += (long double)A[(size_t)i * K + k] * (long double)B[(size_t)j * K + k];
        bound += fabsl((long double)A[(size_t)i * K + k] * (long double)B[(size_t)j * K + k]);
        f64 = fma(A[(size_t)i * K + k], B[(size_t)j * K + k], f64);
      }
      worst_oz = fmaxl(worst_oz, fabsl((long double)C[(size_t)i * N + j] - ref) / bound);
      worst_f64 = fmaxl(worst_f64, fabsl((long double)f64 - ref) / bound);
    }
  printf("Ozaki tile 128 x 64 x %d, %d digit planes (28 INT8 plane products per slab): max |C - ref| / sum|a b| = %.2Le "
         "(plain FP64 FMA loop: %.2Le)\n", K, S, worst_oz, worst_f64);
  return (worst_oz < 1e-11L && mismatches == 0) ? 0 : 1;
}
