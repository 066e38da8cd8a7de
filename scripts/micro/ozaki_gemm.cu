// Micro-probe for DESIGN.md §9 (1), third step: a multi-CTA (one 128 x 64 tile per CTA, UNPIPELINED) FP64 GEMM
// C = A B^T on the INT8 tensor cores by the Ozaki scheme, to see where a first naive version lands against DMMA.
//   kernel 1 (slice_planes_kernel): FP64 operand -> S signed 7-bit digit planes [S][rows][K] (int8) in global memory
//   kernel 2 (ozaki_gemm_kernel)  : per K slab of 128: copy 7 + 7 planes into SWIZZLE_128B shared-memory tiles, issue the
//                                   28 x 4 tcgen05.mma kind::i8 into 7 TMEM accumulators, wait; epilogue recombines in FP64.
// No TMA, no double buffering, no overlap of copy and MMA: every CTA alternates between the two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neo_ls_svm_b200/csrc -o scripts/micro/ozaki_gemm.bin scripts/micro/ozaki_gemm.cu
// Not product code.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include "ptx.cuh"
using namespace nls;

constexpr int BM_ = 128, BN_ = 64, KB = 128, UMMA_K = 32, S = 7, THREADS = 256;

__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// planes[p][r][k] = p-th signed 7-bit digit of X[r][k] 2^-ex[r]   (|X[r][k]| 2^-ex[r] <= 1/2)
__global__ void slice_planes_kernel(const double* __restrict__ X, long long rows, int K, const int* __restrict__ ex,
                                    int8_t* __restrict__ planes) {
  const long long total = rows * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    double x = ldexp(X[e], -ex[e / K]);
#pragma unroll
    for (int p = 0; p < S; ++p) {
      x *= 128.0;
      const int q = __double2int_rn(x);
      x -= (double)q;
      planes[(long long)p * total + e] = (int8_t)q;
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) ozaki_gemm_kernel(const int8_t* __restrict__ Ap, const int8_t* __restrict__ Bp,
                                                                const int* __restrict__ exA, const int* __restrict__ exB,
                                                                double* __restrict__ C, int Mrows, int Ncols, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int PA = BM_ * KB, PB = BN_ * KB;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * PA;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_holder;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM_, n0 = blockIdx.x * BN_;
  constexpr uint32_t COLS = 512;
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM_ >> 4) << 24);
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
  const long long strideA = (long long)Mrows * K, strideB = (long long)Ncols * K;
  for (int k0 = 0; k0 < K; k0 += KB) {
    // 16-byte chunks of the 7 + 7 plane slabs into their swizzled places
    for (int e = tid; e < S * BM_ * 8; e += THREADS) {
      const int p = e / (BM_ * 8), r = (e >> 3) % BM_, c = e & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(Ap + p * strideA + (long long)(m0 + r) * K + k0 + c * 16);
      *reinterpret_cast<uint4*>(sA + p * PA + r * KB + ((c ^ (r & 7)) << 4)) = v;
    }
    for (int e = tid; e < S * BN_ * 8; e += THREADS) {
      const int p = e / (BN_ * 8), r = (e >> 3) % BN_, c = e & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(Bp + p * strideB + (long long)(n0 + r) * K + k0 + c * 16);
      *reinterpret_cast<uint4*>(sB + p * PB + r * KB + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      uint32_t written = k0 > 0 ? 0x7fu : 0u;
      for (int p = 0; p < S; ++p)
        for (int q = 0; p + q < S; ++q) {
          const int t = p + q;
#pragma unroll
          for (int k = 0; k < KB / UMMA_K; ++k) {
            const uint64_t da = umma_desc_k_sw128(smem_u32(sA + p * PA)) + (uint64_t)((k * UMMA_K) >> 4);
            const uint64_t db = umma_desc_k_sw128(smem_u32(sB + q * PB)) + (uint64_t)((k * UMMA_K) >> 4);
            const uint32_t acc = ((written >> t) & 1u) | (k > 0 ? 1u : 0u);
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                    tmem + (uint32_t)(t * BN_)),
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
  // Epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 and the column half w / 4.
  const int row = 32 * (warp & 3) + lane;
  const int chalf = (warp >> 2) * (BN_ / 2);
  for (int c0 = chalf; c0 < chalf + BN_ / 2; c0 += 8) {
    double sum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = 0.0;
    for (int t = S - 1; t >= 0; --t) {
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(t * BN_ + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const double w = ldexp(1.0, -7 * (t + 2));
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] = fma((double)(int32_t)r[j], w, sum[j]);
    }
    const int ea = exA[m0 + row];
#pragma unroll
    for (int j = 0; j < 8; ++j) C[(size_t)(m0 + row) * Ncols + n0 + c0 + j] = ldexp(sum[j], ea + exB[n0 + c0 + j]);
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
    ex[r] = amax > 0.0 ? (int)ceil(log2(amax)) + 1 : 0;
  }
}

int main() {
  // Shape of one projection chunk's real product: 16,384 rows x 1,024 columns x K = 2,048.
  const int Mr = 16384, Nc = 1024, K = 2048;
  std::vector<double> A((size_t)Mr * K), B((size_t)Nc * K);
  srand(5);
  for (auto& v : A) v = (rand() / (double)RAND_MAX - 0.5) * 0.0625;  // like cos/sin / sqrt(D)
  for (auto& v : B) v = (rand() / (double)RAND_MAX - 0.5) * pow(10.0, -3.0 * rand() / (double)RAND_MAX);
  std::vector<int> exA, exB;
  exponents(A, Mr, K, exA);
  exponents(B, Nc, K, exB);
  double *dA, *dB, *dC;
  int *dexA, *dexB;
  int8_t *pA, *pB;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, (size_t)Mr * Nc * 8);
  cudaMalloc(&dexA, Mr * 4); cudaMalloc(&dexB, Nc * 4);
  cudaMalloc(&pA, (size_t)S * Mr * K); cudaMalloc(&pB, (size_t)S * Nc * K);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dexA, exA.data(), Mr * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dexB, exB.data(), Nc * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + (size_t)S * (BM_ + BN_) * KB;
  cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  float ms_slice = 0, ms_gemm = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    slice_planes_kernel<<<148 * 8, 256>>>(dA, Mr, K, dexA, pA);
    slice_planes_kernel<<<148 * 8, 256>>>(dB, Nc, K, dexB, pB);
    cudaEventRecord(e1);
    ozaki_gemm_kernel<<<dim3(Nc / BN_, Mr / BM_), THREADS, smem>>>(pA, pB, dexA, dexB, dC, Mr, Nc, K);
    cudaEventRecord(e2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    cudaEventElapsedTime(&ms_slice, e0, e1);
    cudaEventElapsedTime(&ms_gemm, e1, e2);
  }
  std::vector<double> C((size_t)Mr * Nc);
  cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
  // Check 4,000 random entries: bitwise against the host emulation of the scheme, and the error against long double.
  long long mism = 0;
  long double worst = 0;
  for (int sidx = 0; sidx < 4000; ++sidx) {
    const int i = rand() % Mr, j = rand() % Nc;
    long long acc[S] = {0};
    long double ref = 0, bound = 0;
    std::vector<int8_t> da(S * K), db(S * K);
    for (int k = 0; k < K; ++k) {
      double x = ldexp(A[(size_t)i * K + k], -exA[i]), y = ldexp(B[(size_t)j * K + k], -exB[j]);
      for (int p = 0; p < S; ++p) {
        x *= 128.0; int q = (int)nearbyint(x); x -= q; da[p * K + k] = (int8_t)q;
        y *= 128.0; q = (int)nearbyint(y); y -= q; db[p * K + k] = (int8_t)q;
      }
      ref += (long double)A[(size_t)i * K + k] * (long double)B[(size_t)j * K + k];
      bound += fabsl((long double)A[(size_t)i * K + k] * (long double)B[(size_t)j * K + k]);
    }
    for (int p = 0; p < S; ++p)
      for (int q = 0; p + q < S; ++q) {
        long long d = 0;
        for (int k = 0; k < K; ++k) d += (long long)da[p * K + k] * (long long)db[q * K + k];
        acc[p + q] += d;
      }
    double sum = 0.0;
    for (int t = S - 1; t >= 0; --t) sum = fma((double)acc[t], ldexp(1.0, -7 * (t + 2)), sum);
    if (ldexp(sum, exA[i] + exB[j]) != C[(size_t)i * Nc + j]) ++mism;
    worst = fmaxl(worst, fabsl((long double)C[(size_t)i * Nc + j] - ref) / bound);
  }
  const double flops = 2.0 * Mr * Nc * (double)K;
  printf("Ozaki GEMM %d x %d x %d, %d planes, unpipelined: GEMM kernel %.3f ms = %.1f TFLOP/s FP64-equivalent (slicing %.3f ms; "
         "both %.1f TFLOP/s); 4000 sampled entries: %lld differ from the host emulation, max |C - ref| / sum|a b| = %.2Le\n",
         Mr, Nc, K, S, ms_gemm, flops / (ms_gemm * 1e-3) / 1e12, ms_slice, flops / ((ms_gemm + ms_slice) * 1e-3) / 1e12, mism, worst);
  return mism == 0 ? 0 : 1;
}
