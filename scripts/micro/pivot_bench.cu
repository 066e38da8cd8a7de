// Micro-benchmark / accuracy probe of the 16 x 16 pivot solve of jacobi_wide.cuh (one warp per block).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neo_ls_svm_b200/csrc -o gpurun_out/pivot_bench scripts/micro/pivot_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "jacobi_wide.cuh"
using namespace nls;
constexpr int JBW = 8;
using C = WideCfg<JBW>;

template <int NW>
__global__ void solve_kernel(const double* S_in, double* J_out, double* S_out, long long* cycles, double thr, int cross,
                             int inner, int reps) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  constexpr int P = C::P, SP = C::SP;
  const double* Sb = S_in + (size_t)blockIdx.x * C::JSZ;
  long long t = 0;
  for (int rep = 0; rep < reps; ++rep) {
    for (int e = tid; e < P * P; e += 32 * NW) {
      sm[(e / P) * SP + e % P] = Sb[e];
      sm[P * SP + (e / P) * SP + e % P] = Sb[P * P + e];
    }
    __syncthreads();
    const long long t0 = clock64();
    pivot_rotate_w<JBW, NW>(sm, thr, cross != 0, inner, tid);
    t += clock64() - t0;
    __syncthreads();
  }
  pivot_store_w<JBW, NW>(sm, J_out + (size_t)blockIdx.x * C::JSZ, S_out + (size_t)blockIdx.x * C::JSZ, tid);
  if (tid == 0) cycles[blockIdx.x] = t / reps;
}

__global__ void lat_kernel(double* out, long long* cyc, double x) {
  double y = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    y = fma(y, 1.0000001, 1e-9); y = fma(y, 0.9999999, 1e-9); y = fma(y, 1.0000001, 1e-9); y = fma(y, 0.9999999, 1e-9);
  }
  long long t1 = clock64();
  float f = (float)y;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { f = rsqrtf(f) + 1.0f; f = rsqrtf(f) + 1.0f; f = rsqrtf(f) + 1.0f; f = rsqrtf(f) + 1.0f; }
  long long t2 = clock64();
  double z = y;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { z = (double)(float)z + 1e-9; z = (double)(float)z + 1e-9; z = (double)(float)z + 1e-9; z = (double)(float)z + 1e-9; }
  long long t3 = clock64();
  out[0] = y + f + z;
  cyc[0] = (t1 - t0) / 1024; cyc[1] = (t2 - t1) / 1024; cyc[2] = (t3 - t2) / 1024;
}

int main() {
  const int nblk = 64, P = C::P;
  {
    double* o; long long* c; cudaMalloc(&o, 8); cudaMalloc(&c, 24);
    lat_kernel<<<1, 32>>>(o, c, 1.0);
    long long h[3]; cudaMemcpy(h, c, 24, cudaMemcpyDeviceToHost);
    printf("dependent latency (cycles): DFMA %lld, rsqrtf+FADD %lld, F2F.f32<-f64 + F2F.f64<-f32 + DADD %lld\n", h[0], h[1], h[2]);
  }
  std::vector<double> S((size_t)nblk * C::JSZ);
  srand(1);
  for (int b = 0; b < nblk; ++b) {
    double* Sr = &S[(size_t)b * C::JSZ];
    double* Si = Sr + P * P;
    for (int i = 0; i < P; ++i)
      for (int j = i; j < P; ++j) {
        const double re = rand() / (double)RAND_MAX - 0.5, im = (i == j) ? 0.0 : rand() / (double)RAND_MAX - 0.5;
        const double sc = (i == j) ? 4.0 : 1.0;
        Sr[i * P + j] = sc * re; Si[i * P + j] = im;
        Sr[j * P + i] = sc * re; Si[j * P + i] = -im;
      }
  }
  double *dS, *dJ, *dSo; long long* dC;
  cudaMalloc(&dS, S.size() * 8); cudaMalloc(&dJ, S.size() * 8); cudaMalloc(&dSo, S.size() * 8); cudaMalloc(&dC, nblk * 8);
  cudaMemcpy(dS, S.data(), S.size() * 8, cudaMemcpyHostToDevice);
  for (int nw = 1; nw <= 2; ++nw)
  for (int cross = 0; cross < 2; ++cross)
    for (int inner = 1; inner <= 8; inner *= 8) {
      if (nw == 1) solve_kernel<1><<<nblk, 32, C::PIV_SM * 8>>>(dS, dJ, dSo, dC, 1e-30, cross, inner, 20);
      else solve_kernel<2><<<nblk, 64, C::PIV_SM * 8>>>(dS, dJ, dSo, dC, 1e-30, cross, inner, 20);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<double> J(S.size()), So(S.size()); std::vector<long long> cyc(nblk);
      cudaMemcpy(J.data(), dJ, J.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(So.data(), dSo, So.size() * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(cyc.data(), dC, nblk * 8, cudaMemcpyDeviceToHost);
      double uni = 0, res = 0, off = 0; long long cmax = 0;
      for (int b = 0; b < nblk; ++b) {
        typedef std::complex<double> cd;
        auto Jm = [&](int i, int j) { return cd(J[(size_t)b * C::JSZ + i * P + j], J[(size_t)b * C::JSZ + P * P + i * P + j]); };
        auto Sm = [&](int i, int j) { return cd(S[(size_t)b * C::JSZ + i * P + j], S[(size_t)b * C::JSZ + P * P + i * P + j]); };
        auto Om = [&](int i, int j) { return cd(So[(size_t)b * C::JSZ + i * P + j], So[(size_t)b * C::JSZ + P * P + i * P + j]); };
        for (int i = 0; i < P; ++i)
          for (int j = 0; j < P; ++j) {
            cd u = 0, g = 0;
            for (int k = 0; k < P; ++k) u += std::conj(Jm(k, i)) * Jm(k, j);
            for (int k = 0; k < P; ++k)
              for (int l = 0; l < P; ++l) g += std::conj(Jm(k, i)) * Sm(k, l) * Jm(l, j);
            uni = fmax(uni, std::abs(u - cd(i == j ? 1.0 : 0.0)));
            res = fmax(res, std::abs(g - Om(i, j)));
            if (i != j) off = fmax(off, std::abs(Om(i, j)));
          }
        cmax = cyc[b] > cmax ? cyc[b] : cmax;
      }
      printf("warps=%d cross=%d inner=%d : cycles/solve %lld  |J^H J - I| %.2e  |J^H S J - S_final| %.2e  max offdiag %.2e\n", nw, cross,
             inner, cmax, uni, res, off);
    }
  return 0;
}
