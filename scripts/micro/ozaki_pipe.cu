// Probe for csrc/ozaki.cuh: the pipelined INT8 (Ozaki) GEMM core on one projection chunk's shape.
//   rows x (2 Dp) planar feature chunk  times  the [Re Q^T ; Im Q^T] basis  ->  Re T, Im T  (EpiStore), and the
//   fused P / U epilogue (EpiProject) for timing.
// Checks sampled entries bit for bit against a host emulation of the scheme (exact integer plane sums, same
// recombination order) and reports the error against long double, for both tile images (SWIZZLE_32B and
// no-swizzle), then times the kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I neo_ls_svm_b200/csrc -o scripts/micro/ozaki_pipe.bin scripts/micro/ozaki_pipe.cu
// Not product code.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ozaki.cuh"
using namespace nls;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

struct Problem {
  int rows, D, Dp, Np, m, nks, row_blocks, n_tiles, eA;
  std::vector<double> psi, bt;
  double *d_psi, *d_bt, *d_cs, *d_tr, *d_ti, *d_P, *d_U, *d_vec;
  int* d_ex;
  int8_t *d_a, *d_b;
  std::vector<int> ex;
};

template <int LAYOUT, bool A_TMEM>
static int run(Problem& pr, bool timing) {
  const int sms = 148;
  oz::basis_exponent_kernel<<<(pr.n_tiles * 32 + 7) / 8, 256>>>(pr.d_bt, pr.Np, pr.Dp, pr.D, pr.m, pr.n_tiles * 32, pr.eA, pr.d_ex, pr.d_cs);
  oz::slice_basis_kernel<LAYOUT><<<sms * 8, 256>>>(pr.d_bt, pr.Np, pr.Dp, pr.D, pr.m, pr.d_ex, pr.nks, pr.n_tiles, pr.d_b);
  const double scale = ldexp(1.0, oz::FRAC_BITS - pr.eA);
  oz::slice_rows_kernel<LAYOUT><<<sms * 16, 256>>>(pr.d_psi, 2LL * pr.Dp, pr.rows, pr.D, pr.Dp, scale, pr.nks, pr.row_blocks, pr.d_a);
  CK(cudaDeviceSynchronize());
  oz::GemmParams g{pr.d_a, pr.d_b, pr.nks, pr.row_blocks, pr.n_tiles, 0, pr.row_blocks * pr.n_tiles, 1, pr.nks, 0, 1, 1 << 30, 0};
  const long long ld = pr.Np;
  oz::EpiStore::Params es{pr.rows, pr.m, pr.d_cs, pr.d_tr, pr.d_ti, ld};
  CK(cudaFuncSetAttribute(oz::gemm_kernel_i8<LAYOUT, oz::EpiStore, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
  CK(cudaFuncSetAttribute(oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
  const int grid = (int)std::min<long long>((long long)pr.row_blocks * pr.n_tiles, sms);
  CK(cudaMemset(pr.d_tr, 0, (size_t)pr.rows * ld * 8));
  CK(cudaMemset(pr.d_ti, 0, (size_t)pr.rows * ld * 8));
  if (A_TMEM)
    oz::gemm_kernel_i8<LAYOUT, oz::EpiStore, true><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, es);
  else
    oz::gemm_kernel_i8<LAYOUT, oz::EpiStore, false><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, es);
  CK(cudaDeviceSynchronize());
  std::vector<double> tr((size_t)pr.rows * ld), ti((size_t)pr.rows * ld);
  CK(cudaMemcpy(tr.data(), pr.d_tr, tr.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ti.data(), pr.d_ti, ti.size() * 8, cudaMemcpyDeviceToHost));
  if (pr.ex.empty()) {
    pr.ex.resize(pr.n_tiles * 32);
    CK(cudaMemcpy(pr.ex.data(), pr.d_ex, pr.ex.size() * 4, cudaMemcpyDeviceToHost));
  }
  // host emulation on sampled entries
  const int K = pr.nks * oz::KS;
  long long mism = 0;
  long double worst = 0;
  srand(11);
  const int samples = 3000;
  std::vector<int> da(oz::S * K), dbr(oz::S * K), dbi(oz::S * K);
  for (int sidx = 0; sidx < samples; ++sidx) {
    int i = rand() % pr.rows, j = rand() % pr.m;
    if (sidx < 8) { i = sidx & 1 ? pr.rows - 1 : 0; j = sidx & 2 ? pr.m - 1 : 0; }
    const double sa = ldexp(1.0, oz::FRAC_BITS - pr.eA), sb = ldexp(1.0, oz::FRAC_BITS - pr.ex[j]);
    long double ref_r = 0, ref_i = 0, bound = 0;
    for (int k = 0; k < K; ++k) {
      const int half = (k / oz::KS) & 1, col = (k / oz::KS / 2) * oz::KS + k % oz::KS;  // interleaved K order of ozaki.cuh
      const double a = col < pr.D ? pr.psi[(size_t)i * 2 * pr.Dp + half * pr.Dp + col] : 0.0;
      const double ya = col < pr.D ? pr.bt[(size_t)j * pr.Dp + col] : 0.0, yb = col < pr.D ? pr.bt[(size_t)(pr.Np + j) * pr.Dp + col] : 0.0;
      const double br = half == 0 ? ya : yb, bi = half == 0 ? -yb : ya;
      const unsigned long long qa = oz::digit_bytes(oz::quantise(a, sa)), qr = oz::digit_bytes(oz::quantise(br, sb)), qi = oz::digit_bytes(oz::quantise(bi, sb));
      for (int p = 0; p < oz::S; ++p) {
        da[p * K + k] = oz::digit(qa, p);
        dbr[p * K + k] = oz::digit(qr, p);
        dbi[p * K + k] = oz::digit(qi, p);
      }
      ref_r += (long double)a * br;
      ref_i += (long double)a * bi;
      bound += fabsl((long double)a * br);
    }
    double sr = 0, si = 0;
    for (int t = oz::S - 1; t >= 0; --t) {
      long long ar = 0, ai = 0;
      for (int p = 0; p <= t; ++p)
        for (int k = 0; k < K; ++k) {
          ar += (long long)da[p * K + k] * dbr[(t - p) * K + k];
          ai += (long long)da[p * K + k] * dbi[(t - p) * K + k];
        }
      const double w = oz::level_weight(t);
      sr = fma((double)ar, w, sr);
      si = fma((double)ai, w, si);
    }
    const double cs = ldexp(1.0, pr.eA + pr.ex[j]);
    if (sr * cs != tr[(size_t)i * ld + j] || si * cs != ti[(size_t)i * ld + j]) {
      if (mism < 5) printf("  mismatch (%d,%d): host %.17g %.17g  device %.17g %.17g\n", i, j, sr * cs, si * cs, tr[(size_t)i * ld + j], ti[(size_t)i * ld + j]);
      ++mism;
    }
    worst = fmaxl(worst, fmaxl(fabsl(tr[(size_t)i * ld + j] - ref_r), fabsl(ti[(size_t)i * ld + j] - ref_i)) / bound);
  }
  printf("A from %s, layout %d: %d sampled entries, %lld differ from the host emulation; max |T - ref| / sum|a b| = %.2Le\n", A_TMEM ? "TMEM" : "smem", LAYOUT, samples, mism, worst);
  if (timing) {
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    oz::EpiProject::Params epp{pr.rows, pr.m, pr.d_cs, pr.d_vec, pr.d_vec + pr.Np, pr.d_vec + 2 * pr.Np, pr.d_vec + 3 * pr.Np, 3.0, pr.d_P, pr.d_U, ld};
    float ms_sl = 0, ms_st = 0, ms_pj = 0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      oz::slice_rows_kernel<LAYOUT><<<sms * 16, 256>>>(pr.d_psi, 2LL * pr.Dp, pr.rows, pr.D, pr.Dp, scale, pr.nks, pr.row_blocks, pr.d_a);
      cudaEventRecord(e1);
      oz::gemm_kernel_i8<LAYOUT, oz::EpiStore, A_TMEM><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, es);
      cudaEventRecord(e2);
      oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, A_TMEM><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, epp);
      cudaEventRecord(e3);
      CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms_sl, e0, e1);
      cudaEventElapsedTime(&ms_st, e1, e2);
      cudaEventElapsedTime(&ms_pj, e2, e3);
    }
    if (A_TMEM) {
      float ms_v[4] = {0, 0, 0, 0};
      CK(cudaFuncSetAttribute(oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
      CK(cudaFuncSetAttribute(oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
      CK(cudaFuncSetAttribute(oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
      CK(cudaFuncSetAttribute(oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t ev[5];
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0]);
        oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 2><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, epp);
        cudaEventRecord(ev[1]);
        oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 3><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, epp);
        cudaEventRecord(ev[2]);
        oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 4><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, epp);
        cudaEventRecord(ev[3]);
        oz::gemm_kernel_i8<LAYOUT, oz::EpiProject, true, 5><<<grid, oz::THREADS, oz::SMEM_BYTES>>>(g, epp);
        cudaEventRecord(ev[4]);
        CK(cudaDeviceSynchronize());
        for (int v = 0; v < 4; ++v) cudaEventElapsedTime(&ms_v[v], ev[v], ev[v + 1]);
      }
      printf("  leading planes through TMEM: 2 -> %.3f ms, 3 -> %.3f ms, 4 -> %.3f ms, 5 -> %.3f ms (7 -> %.3f ms)\n", ms_v[0], ms_v[1], ms_v[2], ms_v[3], ms_pj);
    }
    const double flops = 8.0 * pr.rows * (double)pr.D * pr.m;
    printf("A from %s, layout %d: rows %d D %d m %d: slicing %.3f ms (%.0f GB/s), store-epilogue GEMM %.3f ms = %.1f TFLOP/s, project-epilogue GEMM %.3f ms = %.1f TFLOP/s FP64-equivalent (DMMA peak 37.1)\n",
           A_TMEM ? "TMEM" : "smem", LAYOUT, pr.rows, pr.D, pr.m, ms_sl, (pr.rows * 2.0 * pr.Dp * (8 + oz::S)) / ms_sl * 1e-6, ms_st, flops / ms_st * 1e-9, ms_pj, flops / ms_pj * 1e-9);
  }
  return mism == 0 ? 0 : 1;
}

static void make(Problem& pr, int rows, int D, int m) {
  pr.rows = rows; pr.D = D; pr.m = m;
  pr.Dp = (D + 15) / 16 * 16;
  pr.Np = (m + 63) / 64 * 64;
  pr.nks = oz::feature_ksteps(D);
  pr.row_blocks = (rows + oz::TM - 1) / oz::TM;
  pr.n_tiles = (m + 31) / 32;
  const int K = 2 * pr.Dp;
  const double dinv = 1.0 / sqrt((double)D);
  pr.eA = oz::scale_exponent(dinv);
  pr.psi.assign((size_t)rows * K, 0.0);
  pr.bt.assign((size_t)2 * pr.Np * pr.Dp, 0.0);
  srand(5);
  for (int i = 0; i < rows; ++i)
    for (int k = 0; k < D; ++k) {
      const double z = 6.283185307179586 * rand() / (double)RAND_MAX;
      pr.psi[(size_t)i * K + k] = cos(z) * dinv;
      pr.psi[(size_t)i * K + pr.Dp + k] = sin(z) * dinv;
    }
  for (int j = 0; j < m; ++j) {
    const double mag = pow(10.0, -3.0 * rand() / (double)RAND_MAX);
    for (int k = 0; k < D; ++k) {
      pr.bt[(size_t)j * pr.Dp + k] = (rand() / (double)RAND_MAX - 0.5) * mag;
      pr.bt[(size_t)(pr.Np + j) * pr.Dp + k] = (rand() / (double)RAND_MAX - 0.5) * mag;
    }
  }
  CK(cudaMalloc(&pr.d_psi, pr.psi.size() * 8)); CK(cudaMalloc(&pr.d_bt, pr.bt.size() * 8));
  CK(cudaMemcpy(pr.d_psi, pr.psi.data(), pr.psi.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pr.d_bt, pr.bt.data(), pr.bt.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&pr.d_cs, pr.n_tiles * 32 * 8)); CK(cudaMalloc(&pr.d_ex, pr.n_tiles * 32 * 4));
  CK(cudaMalloc(&pr.d_tr, (size_t)rows * pr.Np * 8)); CK(cudaMalloc(&pr.d_ti, (size_t)rows * pr.Np * 8));
  CK(cudaMalloc(&pr.d_P, (size_t)rows * pr.Np * 8)); CK(cudaMalloc(&pr.d_U, (size_t)rows * pr.Np * 8));
  CK(cudaMalloc(&pr.d_vec, (size_t)4 * pr.Np * 8)); CK(cudaMemset(pr.d_vec, 0, (size_t)4 * pr.Np * 8));
  CK(cudaMalloc(&pr.d_a, (size_t)pr.row_blocks * pr.nks * oz::A_STAGE)); CK(cudaMalloc(&pr.d_b, (size_t)pr.n_tiles * pr.nks * oz::B_STAGE));
}

int main(int argc, char** argv) {
  int rc = 0;
  {  // ragged: rows, D and m off every tile boundary
    Problem pr;
    make(pr, 1000, 100, 101);
    rc |= run<6, true>(pr, false);
    rc |= run<6, false>(pr, false) << 1;
  }
  if (argc > 1) {  // "small": the ragged case only (for compute-sanitizer)
    printf("rc = %d\n", rc);
    return rc;
  }
  {  // one projection chunk of C3
    Problem pr;
    make(pr, 32768, 1024, 1024);
    rc |= run<6, true>(pr, true) << 2;
    rc |= run<6, false>(pr, true) << 3;
  }
  printf("rc = %d\n", rc);
  return rc;
}
