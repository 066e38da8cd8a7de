// stedc_cpu.cpp — CPU build of the divide-and-conquer tridiagonal eigensolver's host logic, for the "not gpu" tests.
//
// TEST INFRASTRUCTURE: it drives the SAME headers the CUDA library uses (csrc/stedc_host.h: tree + deflation,
// csrc/secular.h: secular roots) and replaces only the device kernels by plain loops (Jacobi leaves, z-hat, eigenvector
// coefficients, naive GEMM), so that the sequential logic can be checked against LAPACK without a GPU.
// Built by tests/test_stedc_host.py with g++ into a shared object under the pytest tmp dir.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../neo_ls_svm_b200/csrc/secular.h"
#include "../../neo_ls_svm_b200/csrc/stedc_host.h"

using namespace nls;

// Cyclic Jacobi on a small symmetric matrix (what stedc_leaf_kernel does with one warp).
static void leaf_jacobi(int s, std::vector<double>& S, std::vector<double>& V, std::vector<double>& lam) {
  V.assign((size_t)s * s, 0.0);
  for (int i = 0; i < s; ++i) V[(size_t)i * s + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int p = 0; p < s; ++p)
      for (int q = 0; q < s; ++q) (p == q ? dia : off) += S[(size_t)p * s + q] * S[(size_t)p * s + q];
    if (off <= 1e-40 * dia || off == 0.0) break;
    for (int p = 0; p < s - 1; ++p)
      for (int q = p + 1; q < s; ++q) {
        const double apq = S[(size_t)p * s + q];
        if (apq == 0.0) continue;
        const double app = S[(size_t)p * s + p], aqq = S[(size_t)q * s + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < s; ++k) {
          const double skp = S[(size_t)k * s + p], skq = S[(size_t)k * s + q];
          S[(size_t)k * s + p] = c * skp - sn * skq;
          S[(size_t)k * s + q] = sn * skp + c * skq;
        }
        for (int k = 0; k < s; ++k) {
          const double spk = S[(size_t)p * s + k], sqk = S[(size_t)q * s + k];
          S[(size_t)p * s + k] = c * spk - sn * sqk;
          S[(size_t)q * s + k] = sn * spk + c * sqk;
        }
        for (int k = 0; k < s; ++k) {
          const double vkp = V[(size_t)k * s + p], vkq = V[(size_t)k * s + q];
          V[(size_t)k * s + p] = c * vkp - sn * vkq;
          V[(size_t)k * s + q] = sn * vkp + c * vkq;
        }
      }
  }
  lam.resize(s);
  for (int i = 0; i < s; ++i) lam[i] = S[(size_t)i * s + i];
}

extern "C" int stedc_cpu(int n, const double* d_in, const double* e_in, int leaf, double* lam_out, double* Z_out,
                         int* stats /* [2]: merges, total deflated */) {
  std::vector<double> d(d_in, d_in + n), e(e_in, e_in + (n > 1 ? n - 1 : 0));
  double scale = 0.0;
  for (double v : d) scale = std::max(scale, std::fabs(v));
  for (double v : e) scale = std::max(scale, std::fabs(v));
  if (scale == 0.0) scale = 1.0;
  for (double& v : d) v /= scale;
  for (double& v : e) v /= scale;
  std::vector<std::vector<StedcNode>> levels;
  stedc_build_tree(n, leaf, &levels);
  // cut adjustments at every merge boundary
  for (size_t h = 1; h < levels.size(); ++h)
    for (const StedcNode& nd : levels[h]) {
      const double a = std::fabs(e[nd.mid - 1]);
      d[nd.mid - 1] -= a;
      d[nd.mid] -= a;
    }
  std::vector<double> Q((size_t)n * n, 0.0), Qn((size_t)n * n, 0.0), dv(n);
  for (const StedcNode& nd : levels[0]) {
    const int s = nd.hi - nd.lo;
    std::vector<double> S((size_t)s * s, 0.0), V, lam;
    for (int i = 0; i < s; ++i) {
      S[(size_t)i * s + i] = d[nd.lo + i];
      if (i + 1 < s) S[(size_t)i * s + i + 1] = S[(size_t)(i + 1) * s + i] = e[nd.lo + i];
    }
    leaf_jacobi(s, S, V, lam);
    for (int i = 0; i < s; ++i) {
      dv[nd.lo + i] = lam[i];
      for (int k = 0; k < s; ++k) Q[(size_t)(nd.lo + k) * n + nd.lo + i] = V[(size_t)k * s + i];
    }
  }
  stats[0] = stats[1] = 0;
  for (size_t h = 1; h < levels.size(); ++h) {
    std::fill(Qn.begin(), Qn.end(), 0.0);
    std::vector<double> dvn(n);
    for (const StedcNode& nd : levels[h]) {
      const int lo = nd.lo, mid = nd.mid, hi = nd.hi, nm = hi - lo;
      std::vector<double> z(nm);
      for (int j = lo; j < mid; ++j) z[j - lo] = Q[(size_t)(mid - 1) * n + j];
      for (int j = mid; j < hi; ++j) z[j - lo] = Q[(size_t)mid * n + j];
      StedcDeflation df;
      stedc_deflate(lo, mid, hi, dv.data() + lo, z.data(), e[mid - 1], &df);
      stats[0] += 1;
      stats[1] += nm - df.K;
      for (const StedcRotation& r : df.rot)
        for (int row = lo; row < hi; ++row) {
          const double a = Q[(size_t)row * n + r.col_a], b = Q[(size_t)row * n + r.col_b];
          Q[(size_t)row * n + r.col_a] = r.c * a + r.s * b;
          Q[(size_t)row * n + r.col_b] = -r.s * a + r.c * b;
        }
      const int K = df.K;
      if (K > 0) {
        std::vector<double> z2(K), delta((size_t)K * K), zh(K);
        double z2sum = 0.0;
        for (int j = 0; j < K; ++j) {
          z2[j] = df.w[j] * df.w[j];
          z2sum += z2[j];
        }
        for (int i = 0; i < K; ++i) {
          auto eval = [&](int org, double mu) {
            SecularSums s{0, 0, 0, 0};
            for (int j = 0; j < K; ++j) {
              const double del = (df.dlam[j] - df.dlam[org]) - mu;
              const double t = z2[j] / del;
              if (j <= i) { s.psi += t; s.dpsi += t / del; } else { s.phi += t; s.dphi += t / del; }
            }
            s.psi *= df.rho; s.phi *= df.rho; s.dpsi *= df.rho; s.dphi *= df.rho;
            return s;
          };
          int org;
          double mu;
          secular_solve(i, K, df.rho, i + 1 < K ? df.dlam[i + 1] - df.dlam[i] : 0.0, z2sum, z2[K - 1],
                        K > 1 ? df.dlam[K - 1] - df.dlam[K - 2] : 0.0, eval, &org, &mu);
          for (int j = 0; j < K; ++j) delta[(size_t)i * K + j] = (df.dlam[j] - df.dlam[org]) - mu;
          dvn[lo + i] = df.dlam[org] + mu;
        }
        for (int j = 0; j < K; ++j) {  // Gu-Eisenstat: z_j^2 = prod_i (lam_i - d_j) / (rho prod_{i != j} (d_i - d_j))
          double p = -delta[(size_t)j * K + j];
          for (int i = 0; i < K; ++i)
            if (i != j) p *= -delta[(size_t)i * K + j] / (df.dlam[i] - df.dlam[j]);
          zh[j] = std::copysign(std::sqrt(std::fabs(p) / df.rho), df.w[j]);
        }
        const int kt = df.k1 + df.k3, kb = df.k3 + df.k2;
        std::vector<double> Ut((size_t)K * std::max(kt, 1), 0.0), Ub((size_t)K * std::max(kb, 1), 0.0);
        for (int i = 0; i < K; ++i) {
          double nrm = 0.0;
          for (int j = 0; j < K; ++j) {
            const double u = zh[j] / delta[(size_t)i * K + j];
            nrm += u * u;
          }
          nrm = 1.0 / std::sqrt(nrm);
          for (int j = 0; j < K; ++j) {
            const double u = zh[j] / delta[(size_t)i * K + j] * nrm;
            if (df.pos_top[j] >= 0) Ut[(size_t)i * kt + df.pos_top[j]] = u;
            if (df.pos_bot[j] >= 0) Ub[(size_t)i * kb + df.pos_bot[j]] = u;
          }
        }
        for (int row = lo; row < hi; ++row) {
          const bool top = row < mid;
          const int kk = top ? kt : kb;
          const std::vector<int>& src = top ? df.src_top : df.src_bot;
          const std::vector<double>& U = top ? Ut : Ub;
          for (int i = 0; i < K; ++i) {
            double acc = 0.0;
            for (int q = 0; q < kk; ++q) acc += Q[(size_t)row * n + src[q]] * U[(size_t)i * kk + q];
            Qn[(size_t)row * n + lo + i] = acc;
          }
        }
      }
      for (size_t t = 0; t < df.defl_col.size(); ++t) {
        dvn[lo + K + t] = df.defl_val[t];
        for (int row = lo; row < hi; ++row) Qn[(size_t)row * n + lo + K + t] = Q[(size_t)row * n + df.defl_col[t]];
      }
    }
    Q.swap(Qn);
    dv.swap(dvn);
  }
  std::vector<int> perm(n);
  for (int j = 0; j < n; ++j) perm[j] = j;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return dv[a] < dv[b]; });
  for (int k = 0; k < n; ++k) {
    lam_out[k] = dv[perm[k]] * scale;
    for (int r = 0; r < n; ++r) Z_out[(size_t)r * n + k] = Q[(size_t)r * n + perm[k]];
  }
  return 0;
}
