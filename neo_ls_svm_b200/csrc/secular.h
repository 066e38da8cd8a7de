// secular.h — one root of the secular equation of a rank-one modified diagonal eigenproblem,
//     f(lam) = 1 + rho * sum_j z_j^2 / (d_j - lam) = 0,      d_0 < d_1 < ... < d_{K-1},  rho > 0,
// the inner kernel of the divide-and-conquer tridiagonal eigensolver (csrc/stedc.cuh) that replaces LAPACK's
// zheevr/dsyevd behind scipy.linalg.eigh / np.linalg.eigh at the reference's _neo_ls_svm.py:120 and :265.
//
// Shared by the CUDA kernel (a warp evaluates the sums cooperatively) and by the host build used in the CPU tests
// (tests/native/stedc_cpu.cpp), through the `Eval` functor.  The root is returned as (origin, mu) with
// lam = d[origin] + mu, origin the nearer of the two enclosing poles, so that the differences
// d_j - lam = (d_j - d[origin]) - mu keep full relative accuracy (what the Gu-Eisenstat eigenvector formula needs).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define NLS_HD __host__ __device__ __forceinline__
#else
#define NLS_HD inline
#endif

namespace nls {

struct SecularSums {
  double psi, phi;    // rho * sum_{j <= i} z_j^2 / delta_j  and  rho * sum_{j > i} ...
  double dpsi, dphi;  // derivatives (sums of z_j^2 / delta_j^2, times rho), both > 0
};

// Root of  c + b1 / (D1 - eta) + b2 / (D2 - eta) = 0  that lies in (lo, hi); D1 < 0 < D2 are the distances of the two
// interpolation poles from the current iterate.  Falls back to the bracket midpoint when no root qualifies.
NLS_HD double secular_two_pole_step(double c, double b1, double D1, double b2, double D2, double lo, double hi) {
  const double A = c;
  const double B = -(c * (D1 + D2) + b1 + b2);
  const double C = c * D1 * D2 + b1 * D2 + b2 * D1;
  double disc = B * B - 4.0 * A * C;
  if (disc < 0.0) disc = 0.0;
  const double sq = sqrt(disc);
  const double q = (B >= 0.0) ? -0.5 * (B + sq) : -0.5 * (B - sq);
  double r1 = (A != 0.0) ? q / A : INFINITY;
  double r2 = (q != 0.0) ? C / q : INFINITY;
  if (r1 > lo && r1 < hi) return r1;
  if (r2 > lo && r2 < hi) return r2;
  return 0.5 * (lo + hi);
}

// eval(origin, mu) must return the SecularSums at lam = d[origin] + mu, split at index i.
// dd_i1 = d[i+1] - d[i] (ignored for the last root); z2sum = sum_j z_j^2; z2last, dgap_last = d[K-1] - d[K-2] are
// only used for the last root (i == K - 1).
template <class Eval>
NLS_HD void secular_solve(int i, int K, double rho, double dd_i1, double z2sum, double z2last, double dgap_last,
                          Eval&& eval, int* origin_out, double* mu_out) {
  const double eps = 2.220446049250313e-16;
  const bool last = (i == K - 1);
  int org;
  double lo, hi, mu;
  if (!last) {
    const double mid = 0.5 * dd_i1;
    const SecularSums s = eval(i, mid);
    if (1.0 + s.psi + s.phi > 0.0) {  // root in the left half of the interval: measure from d_i
      org = i; lo = 0.0; hi = mid; mu = 0.5 * mid;
    } else {
      org = i + 1; lo = -mid; hi = 0.0; mu = -0.5 * mid;
    }
  } else {
    org = K - 1; lo = 0.0; hi = rho * z2sum; mu = 0.5 * hi;
    if (K == 1) {
      *origin_out = 0;
      *mu_out = hi;
      return;
    }
  }
  for (int it = 0; it < 200; ++it) {
    const SecularSums s = eval(org, mu);
    const double f = 1.0 + s.psi + s.phi;
    if (f > 0.0) { if (mu < hi) hi = mu; } else { if (mu > lo) lo = mu; }
    const double tol = eps * (8.0 * (fabs(s.psi) + fabs(s.phi)) + 2.0 + 3.0 * fabs(mu) * (s.dpsi + s.dphi));
    const double wid = fmax(fabs(lo), fabs(hi));
    if (fabs(f) <= tol || hi - lo <= 2.0 * eps * wid) break;
    double next;
    if (it >= 60) {
      next = 0.5 * (lo + hi);
    } else if (!last) {
      // psi ~ a1 + b1 / (D1 - eta) around pole i, phi ~ a2 + b2 / (D2 - eta) around pole i + 1 (Bunch-Nielsen-Sorensen)
      const double D1 = (org == i ? 0.0 : -dd_i1) - mu;
      const double D2 = (org == i ? dd_i1 : 0.0) - mu;
      const double b1 = D1 * D1 * s.dpsi, a1 = s.psi - D1 * s.dpsi;
      const double b2 = D2 * D2 * s.dphi, a2 = s.phi - D2 * s.dphi;
      next = mu + secular_two_pole_step(1.0 + a1 + a2, b1, D1, b2, D2, lo - mu, hi - mu);
    } else {
      // last root: the pole K-1 term is kept exactly, the rest of psi is interpolated around pole K-2
      const double D2 = -mu;               // distance to pole K-1 (negative: lam lies beyond it)
      const double D1 = -dgap_last - mu;   // distance to pole K-2
      const double tl = rho * z2last;
      const double psi0 = s.psi - tl / D2, dpsi0 = s.dpsi - tl / (D2 * D2);
      const double b1 = D1 * D1 * dpsi0, a1 = psi0 - D1 * dpsi0;
      next = mu + secular_two_pole_step(1.0 + a1, b1, D1, tl, D2, lo - mu, hi - mu);
    }
    if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);
    if (next == mu) break;
    mu = next;
  }
  *origin_out = org;
  *mu_out = mu;
}

}  // namespace nls
