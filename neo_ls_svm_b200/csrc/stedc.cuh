// stedc.cuh — device kernels of the divide-and-conquer eigensolver for real symmetric tridiagonal matrices, the
// middle stage of the hand-written dense eigensolver (tridiagonalisation -> this -> back-transformation) that replaces
// LAPACK behind scipy.linalg.eigh (_neo_ls_svm.py:120) and np.linalg.eigh (:265).
//
// Division of labour per merge (csrc/stedc_host.h has the host half):
//   host   : sort the children's eigenvalues, deflate (O(n) scans), lay out the operands
//   device : Givens rotations of deflated pairs, column gather, K secular roots (one warp each, all K^2 pole
//            differences kept), Gu-Eisenstat z-hat (K^2 products), normalised eigenvector coefficients, and the
//            two GEMMs  Q_new[top rows] = Q[top, types 1|3] U_top^T,  Q_new[bottom rows] = Q[bottom, types 3|2] U_bot^T
//            on the FP64 DMMA GEMM core (gemm_kernel<MODE_REAL, OpStore<false>>).
// All merges of one tree level are independent and are processed by batched launches (blockIdx.y = merge).
// Every reduction has a fixed order: two runs give bitwise identical results.
#pragma once
#include <cuda_runtime.h>

#include "secular.h"

namespace nls {

struct StedcMergeDesc {
  int lo, mid, hi;
  int K, ktop, kbot;
  int nrot, rot_off;
  int ndefl, pad;
  double rho;
};

struct StedcRotDev {
  int col_a, col_b;
  double c, s;
};

constexpr int STEDC_LEAF = 32;

// ---------------------------------------------------------------------------------------------
// Leaves: cyclic Jacobi on the dense s x s block (s <= 32), one warp per leaf.  Writes the ascending eigenvalues to
// dvals[lo .. hi) and the eigenvectors to the diagonal block Q[lo:hi, lo:hi].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) stedc_leaf_kernel(const double* __restrict__ d, const double* __restrict__ e,
                                                        const int2* __restrict__ leaves, int nleaves,
                                                        double* __restrict__ Q, long long ld,
                                                        double* __restrict__ dvals) {
  __shared__ double Ssm[2][32][33];
  __shared__ double Vsm[2][32][33];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = blockIdx.x * 2 + wib;
  if (leaf >= nleaves) return;
  const int lo = leaves[leaf].x, s = leaves[leaf].y - lo;
  double(*S)[33] = Ssm[wib];
  double(*V)[33] = Vsm[wib];
  for (int r = 0; r < 32; ++r) {
    S[r][lane] = 0.0;
    V[r][lane] = (r == lane) ? 1.0 : 0.0;
  }
  __syncwarp();
  if (lane < s) {
    S[lane][lane] = d[lo + lane];
    if (lane + 1 < s) {
      const double ev = e[lo + lane];
      S[lane][lane + 1] = ev;
      S[lane + 1][lane] = ev;
    }
  }
  __syncwarp();
  double fro = 0.0;
  for (int r = 0; r < s; ++r) fro += S[r][lane] * S[r][lane];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, off);
  const double thr = 3.5e-18 * sqrt(fro);  // eps/64 * ||S||_F: absolute accuracy far below the merge tolerance
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < s - 1; ++p)
      for (int q = p + 1; q < s; ++q) {
        const double apq = S[p][q];
        if (fabs(apq) <= thr) continue;  // uniform over the warp: every lane reads the same entry
        rotated = true;
        const double app = S[p][p], aqq = S[q][q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        __syncwarp();
        {
          const double skp = S[lane][p], skq = S[lane][q];
          S[lane][p] = c * skp - sn * skq;
          S[lane][q] = sn * skp + c * skq;
          const double vkp = V[lane][p], vkq = V[lane][q];
          V[lane][p] = c * vkp - sn * vkq;
          V[lane][q] = sn * vkp + c * vkq;
        }
        __syncwarp();
        {
          const double spk = S[p][lane], sqk = S[q][lane];
          S[p][lane] = c * spk - sn * sqk;
          S[q][lane] = sn * spk + c * sqk;
        }
        __syncwarp();
      }
    if (!rotated) break;
  }
  __syncwarp();
  const double mine = lane < s ? S[lane][lane] : 0.0;
  int rank = 0;
  for (int j = 0; j < s; ++j) {
    const double other = S[j][j];
    rank += (other < mine || (other == mine && j < lane)) ? 1 : 0;
  }
  if (lane < s) {
    dvals[lo + rank] = mine;
    for (int r = 0; r < s; ++r) Q[(long long)(lo + r) * ld + lo + rank] = V[r][lane];
  }
}

// z of every merge of a level: [last row of the top child's Q | first row of the bottom child's Q].
__global__ void stedc_zrows_kernel(const double* __restrict__ Q, long long ld, const StedcMergeDesc* __restrict__ desc,
                                   double* __restrict__ z) {
  const StedcMergeDesc m = desc[blockIdx.y];
  const int j = m.lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m.hi) return;
  z[j] = j < m.mid ? Q[(long long)(m.mid - 1) * ld + j] : Q[(long long)m.mid * ld + j];
}

// Deflation rotations, in sequence, on the rows of the merged block (rows are independent).
__global__ void stedc_rotate_kernel(double* __restrict__ Q, long long ld, const StedcMergeDesc* __restrict__ desc,
                                    const StedcRotDev* __restrict__ rots) {
  const StedcMergeDesc m = desc[blockIdx.y];
  if (m.nrot == 0) return;
  const int row = m.lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m.hi) return;
  double* q = Q + (long long)row * ld;
  for (int r = 0; r < m.nrot; ++r) {
    const StedcRotDev g = rots[m.rot_off + r];
    const double a = q[g.col_a], b = q[g.col_b];
    q[g.col_a] = g.c * a + g.s * b;
    q[g.col_b] = -g.s * a + g.c * b;
  }
}

// Row r of the merged block: Qp[r][0:kdim) = Q[r][src[...]] (top rows take the [type 1 | type 3] columns, bottom rows
// [type 3 | type 2]); deflated columns go straight to their final place in Qnext, their eigenvalues to dnext.
__global__ void __launch_bounds__(128) stedc_gather_kernel(const double* __restrict__ Q, double* __restrict__ Qp,
                                                           double* __restrict__ Qnext, long long ld,
                                                           const StedcMergeDesc* __restrict__ desc,
                                                           const int* __restrict__ src_top,
                                                           const int* __restrict__ src_bot,
                                                           const int* __restrict__ defl_src,
                                                           const double* __restrict__ defl_val,
                                                           double* __restrict__ dnext) {
  const StedcMergeDesc m = desc[blockIdx.y];
  const int row = m.lo + blockIdx.x;
  if (row >= m.hi) return;
  const bool top = row < m.mid;
  const int kdim = top ? m.ktop : m.kbot;
  const int* src = (top ? src_top : src_bot) + m.lo;
  const double* q = Q + (long long)row * ld;
  double* qp = Qp + (long long)row * ld;
  for (int k = threadIdx.x; k < kdim; k += blockDim.x) qp[k] = q[src[k]];
  double* qn = Qnext + (long long)row * ld + m.lo + m.K;
  for (int t = threadIdx.x; t < m.ndefl; t += blockDim.x) qn[t] = q[defl_src[m.lo + m.K + t]];
  if (row == m.lo)
    for (int t = threadIdx.x; t < m.ndefl; t += blockDim.x) dnext[m.lo + m.K + t] = defl_val[m.lo + m.K + t];
}

// Sums of the secular function, evaluated by one warp (lanes stride over the poles, butterfly reduction: every lane
// ends up with the same bits).
struct WarpSecularEval {
  const double* dl;
  const double* z2;
  int K, i, lane;
  double rho;
  NLS_HD SecularSums operator()(int org, double mu) const {
    SecularSums s{0.0, 0.0, 0.0, 0.0};
#ifdef __CUDA_ARCH__
    const double dorg = dl[org];
    for (int j = lane; j < K; j += 32) {
      const double del = (dl[j] - dorg) - mu;
      const double t = z2[j] / del;
      if (j <= i) {
        s.psi += t;
        s.dpsi += t / del;
      } else {
        s.phi += t;
        s.dphi += t / del;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s.psi += __shfl_xor_sync(0xffffffffu, s.psi, off);
      s.phi += __shfl_xor_sync(0xffffffffu, s.phi, off);
      s.dpsi += __shfl_xor_sync(0xffffffffu, s.dpsi, off);
      s.dphi += __shfl_xor_sync(0xffffffffu, s.dphi, off);
    }
    s.psi *= rho;
    s.phi *= rho;
    s.dpsi *= rho;
    s.dphi *= rho;
#endif
    return s;
  }
};

// One warp per secular root: lam_i -> dnext[lo + i], pole differences d_j - lam_i -> Delta[lo + i][j].
// z2 (scratch, n): squares of the kept z components, written by the same kernel's first phase... kept separate
// (stedc_square_kernel) so that this kernel has no intra-grid dependency.
__global__ void stedc_square_kernel(const double* __restrict__ w, const StedcMergeDesc* __restrict__ desc,
                                    double* __restrict__ z2) {
  const StedcMergeDesc m = desc[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m.K) z2[m.lo + j] = w[m.lo + j] * w[m.lo + j];
}

__global__ void __launch_bounds__(256) stedc_secular_kernel(const StedcMergeDesc* __restrict__ desc,
                                                            const double* __restrict__ dlam,
                                                            const double* __restrict__ z2g,
                                                            double* __restrict__ Delta, long long ld,
                                                            double* __restrict__ dnext) {
  const StedcMergeDesc m = desc[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= m.K) return;
  const int K = m.K;
  const double* dl = dlam + m.lo;
  const double* z2 = z2g + m.lo;
  double z2sum = 0.0;
  for (int j = lane; j < K; j += 32) z2sum += z2[j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) z2sum += __shfl_xor_sync(0xffffffffu, z2sum, off);
  WarpSecularEval eval{dl, z2, K, i, lane, m.rho};
  int org;
  double mu;
  secular_solve(i, K, m.rho, i + 1 < K ? dl[i + 1] - dl[i] : 0.0, z2sum, z2[K - 1], K > 1 ? dl[K - 1] - dl[K - 2] : 0.0,
                eval, &org, &mu);
  const double dorg = dl[org];
  double* row = Delta + (long long)(m.lo + i) * ld;
  for (int j = lane; j < K; j += 32) row[j] = (dl[j] - dorg) - mu;
  if (lane == 0) dnext[m.lo + i] = dorg + mu;
}

// Gu-Eisenstat: zhat_j^2 = prod_i (lam_i - d_j) / (rho prod_{i != j} (d_i - d_j)), sign of the original z_j.
// Block (32, 8): 32 consecutive j (coalesced reads of Delta rows), 8 interleaved partial products over i combined in
// a fixed order.
__global__ void __launch_bounds__(256) stedc_zhat_kernel(const StedcMergeDesc* __restrict__ desc,
                                                         const double* __restrict__ dlam, const double* __restrict__ w,
                                                         const double* __restrict__ Delta, long long ld,
                                                         double* __restrict__ zhat) {
  __shared__ double part[8][32];
  const StedcMergeDesc m = desc[blockIdx.y];
  const int K = m.K;
  if (blockIdx.x * 32 >= K) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const double* dl = dlam + m.lo;
  double p = 1.0;
  if (j < K) {
    const double dj = dl[j];
    for (int i = ty; i < K; i += 8) {
      const double del = Delta[(long long)(m.lo + i) * ld + j];
      p *= (i == j) ? -del : (-del) / (dl[i] - dj);
    }
  }
  part[ty][tx] = p;
  __syncthreads();
  if (ty == 0 && j < K) {
    double prod = part[0][tx];
#pragma unroll
    for (int q = 1; q < 8; ++q) prod *= part[q][tx];
    zhat[m.lo + j] = copysign(sqrt(fabs(prod) / m.rho), w[m.lo + j]);
  }
}

// Eigenvector coefficients of root i: u_j = zhat_j / (d_j - lam_i), normalised, scattered into the two GEMM operands
// U_top[lo + i][pos_top[j]] and U_bot[lo + i][pos_bot[j]] (K-contiguous rows).  One block per root.
__global__ void __launch_bounds__(256) stedc_vectors_kernel(const StedcMergeDesc* __restrict__ desc,
                                                            const double* __restrict__ Delta, long long ld,
                                                            const double* __restrict__ zhat,
                                                            const int* __restrict__ pos_top,
                                                            const int* __restrict__ pos_bot,
                                                            double* __restrict__ Utop, double* __restrict__ Ubot) {
  __shared__ double red[256];
  const StedcMergeDesc m = desc[blockIdx.y];
  const int i = blockIdx.x;
  if (i >= m.K) return;
  const int K = m.K;
  const double* del = Delta + (long long)(m.lo + i) * ld;
  const double* zh = zhat + m.lo;
  double acc = 0.0;
  for (int j = threadIdx.x; j < K; j += 256) {
    const double u = zh[j] / del[j];
    acc += u * u;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  const double inv = 1.0 / sqrt(red[0]);
  double* ut = Utop + (long long)(m.lo + i) * ld;
  double* ub = Ubot + (long long)(m.lo + i) * ld;
  for (int j = threadIdx.x; j < K; j += 256) {
    const double u = zh[j] / del[j] * inv;
    const int pt = pos_top[m.lo + j], pb = pos_bot[m.lo + j];
    if (pt >= 0) ut[pt] = u;
    if (pb >= 0) ub[pb] = u;
  }
}

// out[rows x cols] = 0 (a merge whose kept columns all live in the other half).
__global__ void stedc_zero_block_kernel(double* __restrict__ out, long long ld, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x)
    out[(t / cols) * ld + t % cols] = 0.0;
}

// The two off-diagonal blocks Q[lo:mid, mid:hi] and Q[mid:hi, lo:mid] of every merge of the NEXT level: the merges of
// the current level only write their own diagonal blocks of the output buffer, and the next level's rotations and
// deflated-column copies read whole columns of the merged block, so the part outside the children must be zero.
__global__ void stedc_zero_cross_kernel(double* __restrict__ Q, long long ld, const int4* __restrict__ parents) {
  const int4 p = parents[blockIdx.y];  // lo, mid, hi
  const int lo = p.x, mid = p.y, hi = p.z;
  for (int row = lo + blockIdx.x; row < hi; row += gridDim.x) {
    const int c0 = row < mid ? mid : lo, c1 = row < mid ? hi : mid;
    double* q = Q + (long long)row * ld;
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) q[c] = 0.0;
  }
}

// Final ordering + transposition: Yt[k][r] = Q[r][perm[k]] (row k = eigenvector of the k-th smallest eigenvalue),
// lam[k] = dvals[perm[k]] * scale.  32 x 32 tiles through shared memory (coalesced on both sides).
__global__ void __launch_bounds__(256) stedc_finish_kernel(const double* __restrict__ Q, long long ld, int n,
                                                           const int* __restrict__ perm,
                                                           const double* __restrict__ dvals, double scale,
                                                           double* __restrict__ Yt, long long ldy,
                                                           double* __restrict__ lam) {
  __shared__ double tile[32][33];
  const int k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = r0 + rr, k = k0 + tx;
    tile[rr][tx] = (r < n && k < n) ? Q[(long long)r * ld + perm[k]] : 0.0;
  }
  __syncthreads();
  for (int kk = ty; kk < 32; kk += 8) {
    const int k = k0 + kk, r = r0 + tx;
    if (k < n && r < n) Yt[(long long)k * ldy + r] = tile[tx][kk];
  }
  if (blockIdx.y == 0 && ty == 0 && k0 + tx < n) lam[k0 + tx] = dvals[perm[k0 + tx]] * scale;
}

}  // namespace nls
