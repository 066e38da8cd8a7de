// stedc_host.h — host-side logic of the divide-and-conquer tridiagonal eigensolver: the merge tree and the
// deflation step of every merge (the O(n) sequential scans; everything O(n^2) and O(n^3) runs on the device, see
// csrc/stedc.cuh).  Plain C++ with no CUDA dependency so that the CPU tests can drive it (tests/native/stedc_cpu.cpp).
//
// One merge glues the eigen-decompositions of two adjacent diagonal blocks T1', T2' (their coupling entry e removed
// and |e| subtracted from the two neighbouring diagonal entries) into that of
//     blkdiag(T1', T2') + |e| u u^T = Q (D + rho z z^T) Q^T,   z = Q^T u / sqrt(2),  rho = 2 |e|,
// where u has 1 at the last row of block 1 and sign(e) at the first row of block 2.  Deflation removes the components
// of z that are negligible and pairs of (numerically) equal diagonal entries (one Givens rotation each), exactly as
// LAPACK's dlaed2 does; what is left is the K x K secular problem solved on the device.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace nls {

struct StedcNode {
  int lo, mid, hi;   // rows/columns [lo, hi); children [lo, mid) and [mid, hi).  Leaf: mid == -1.
  int height;        // 0 for leaves
};

// Uniform-depth binary tree: every leaf has at most `leaf` entries and all leaves sit at the same depth, so that all
// merges of one height are independent and the eigenvector matrix can be double-buffered level by level.
inline void stedc_build_tree(int n, int leaf, std::vector<std::vector<StedcNode>>* levels) {
  int depth = 0;
  while (((n + (1 << depth) - 1) >> depth) > leaf) ++depth;
  levels->assign(depth + 1, {});
  // ranges at depth k: split points i*n / 2^k
  for (int h = 0; h <= depth; ++h) {
    const int k = depth - h;  // tree depth of the nodes with height h
    const long long parts = 1LL << k;
    for (long long i = 0; i < parts; ++i) {
      StedcNode nd;
      nd.lo = (int)(i * n / parts);
      nd.hi = (int)((i + 1) * n / parts);
      nd.mid = h == 0 ? -1 : (int)((2 * i + 1) * n / (2 * parts));
      nd.height = h;
      (*levels)[h].push_back(nd);
    }
  }
}

struct StedcRotation {
  int col_a, col_b;  // columns (global indices) rotated: a' = c a + s b, b' = -s a + c b
  double c, s;
};

// Result of the deflation of one merge.  All column indices are GLOBAL (lo-based offsets already added).
struct StedcDeflation {
  int K = 0;                          // size of the secular problem
  double rho = 0.0;
  std::vector<double> dlam, w;        // K ascending poles and the matching z components
  std::vector<int> keep_col;          // K source columns, same order as dlam
  std::vector<int> keep_type;         // 1: non-zero in the top block only, 2: bottom only, 3: dense
  std::vector<StedcRotation> rot;     // applied in sequence BEFORE the columns are gathered
  std::vector<int> defl_col;          // source columns of the deflated eigenpairs
  std::vector<double> defl_val;       // their eigenvalues
  // Grouped order [type 1 | type 3 | type 2] of the kept columns: the top block multiplies [1|3], the bottom [3|2].
  int k1 = 0, k2 = 0, k3 = 0;
  std::vector<int> pos_top, pos_bot;  // per kept column (dlam order): its slot in the top / bottom operand, or -1
  std::vector<int> src_top, src_bot;  // per slot: the source column
};

// d: current eigenvalues of the two children, d[j] belongs to column lo + j (any order within a child);
// z: [last row of Q1 | first row of Q2] (unscaled, unsigned); e: the coupling entry.
inline void stedc_deflate(int lo, int mid, int hi, const double* d_in, const double* z_in, double e, StedcDeflation* out) {
  const double eps = 2.220446049250313e-16;
  const int n = hi - lo, n1 = mid - lo;
  std::vector<double> d(d_in, d_in + n), z(z_in, z_in + n);
  const double sgn = e < 0.0 ? -1.0 : 1.0;
  const double inv_sqrt2 = 0.70710678118654752440;
  for (int j = 0; j < n; ++j) z[j] *= (j >= n1 ? sgn : 1.0) * inv_sqrt2;
  const double rho = 2.0 * std::fabs(e);
  std::vector<int> order(n);
  for (int j = 0; j < n; ++j) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
  double dmax = 0.0, zmax = 0.0;
  for (int j = 0; j < n; ++j) {
    dmax = std::max(dmax, std::fabs(d[j]));
    zmax = std::max(zmax, std::fabs(z[j]));
  }
  const double tol = 8.0 * eps * std::max(dmax, zmax);
  std::vector<int> type(n);
  for (int j = 0; j < n; ++j) type[j] = j < n1 ? 1 : 2;
  *out = StedcDeflation();
  out->rho = rho;
  std::vector<int> keep;
  if (rho * zmax <= tol) {
    for (int j : order) {
      out->defl_col.push_back(lo + j);
      out->defl_val.push_back(d[j]);
    }
  } else {
    int pj = -1;
    for (int j : order) {
      if (rho * std::fabs(z[j]) <= tol) {  // negligible coupling: (d_j, column j) is already an eigenpair
        out->defl_col.push_back(lo + j);
        out->defl_val.push_back(d[j]);
        continue;
      }
      if (pj < 0) {
        pj = j;
        continue;
      }
      double s = z[pj], c = z[j];
      const double tau = std::hypot(c, s);
      const double t = d[j] - d[pj];
      c /= tau;
      s = -s / tau;
      if (std::fabs(t * c * s) <= tol) {  // (numerically) equal poles: rotate z_pj into z_j, column pj deflates
        z[j] = tau;
        z[pj] = 0.0;
        out->rot.push_back(StedcRotation{lo + pj, lo + j, c, s});
        if (type[j] != type[pj]) type[j] = 3;
        const double t2 = d[pj] * c * c + d[j] * s * s;
        d[j] = d[pj] * s * s + d[j] * c * c;
        d[pj] = t2;
        out->defl_col.push_back(lo + pj);
        out->defl_val.push_back(d[pj]);
        pj = j;
      } else {
        keep.push_back(pj);
        pj = j;
      }
    }
    if (pj >= 0) keep.push_back(pj);
  }
  const int K = (int)keep.size();
  out->K = K;
  out->dlam.resize(K);
  out->w.resize(K);
  out->keep_col.resize(K);
  out->keep_type.resize(K);
  out->pos_top.assign(K, -1);
  out->pos_bot.assign(K, -1);
  for (int i = 0; i < K; ++i) {
    out->dlam[i] = d[keep[i]];
    out->w[i] = z[keep[i]];
    out->keep_col[i] = lo + keep[i];
    out->keep_type[i] = type[keep[i]];
    out->k1 += type[keep[i]] == 1;
    out->k2 += type[keep[i]] == 2;
    out->k3 += type[keep[i]] == 3;
  }
  // The poles must be strictly increasing for the secular solver; equal neighbours would have been deflated above
  // unless tol is zero (all-zero input), in which case rho * zmax <= tol caught it.
  out->src_top.assign(out->k1 + out->k3, 0);
  out->src_bot.assign(out->k3 + out->k2, 0);
  int c1 = 0, c3 = 0, c2 = 0;
  for (int i = 0; i < K; ++i) {
    const int col = out->keep_col[i];
    if (out->keep_type[i] == 1) {
      out->pos_top[i] = c1;
      out->src_top[c1++] = col;
    } else if (out->keep_type[i] == 3) {
      out->pos_top[i] = out->k1 + c3;
      out->pos_bot[i] = c3;
      out->src_top[out->k1 + c3] = col;
      out->src_bot[c3] = col;
      ++c3;
    } else {
      out->pos_bot[i] = out->k3 + c2;
      out->src_bot[out->k3 + c2] = col;
      ++c2;
    }
  }
}

}  // namespace nls
