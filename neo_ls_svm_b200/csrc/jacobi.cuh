// jacobi.cuh — hand-written Hermitian eigensolver for stage 3, 4-wide-block variant (comparator and fallback of the
// default 8-wide solver in jacobi_wide.cuh; NLS_JACOBI_JB=4 selects it).
//
// Replaces scipy.linalg.eigh at _neo_ls_svm.py:120 for the m x m complex Hermitian matrix A/c.
// (One-sided / Hestenes Jacobi on A was prototyped first and rejected: it implicitly works on A^2,
// and the Gram matrices of this path have spectra graded over >14 decades, on which it stalls.)
//
// The index range [0, mp) (mp = m padded to an even number of 4-wide blocks) is cut into 4-wide blocks;
// a round-robin tournament pairs the blocks so that every round holds nb/2 disjoint pivot pairs of 8
// indices.  A round is either two kernels ("split"):
//   jacobi_pivot_kernel  : one warp per pair diagonalises its 8 x 8 Hermitian pivot block with cyclic
//                          Jacobi rotations in shared memory (32 lanes = 8 rows x 4 disjoint rotations),
//                          accumulating the 8 x 8 unitary J_i.
//   jacobi_update_kernel : G <- J^H G J tile by tile (tile (i,j) = J_i^H G[I_i, I_j] J_j, read and written
//                          by one warp only, so the update is in place) and V <- V J, as complex 8x8x8
//                          products on DMMA.8x8x4.
// or one ("fused", jacobi_round_kernel: the tiles that hold next round's pivots are updated first and published
// through a counter, the pivot warps wait for it and solve while the other warps finish the update).
// A rotation is skipped when |g_pq| <= eps ||G||_F (absolute threshold: the decomposition is backward
// stable w.r.t. ||A||, like LAPACK's zheevr); the sweep loop ends when no pair rotated.  Pad
// rows/columns are exactly zero, never rotate, and are dropped at the end.  One sweep is captured once
// into a CUDA graph and replayed.  A persistent cooperative kernel (software grid barrier) and programmatic
// dependent launch were also built and measured slower than the graph of short kernels (73 and 83 ms against 70 ms
// at m = 1025; profiles/r1_eigensolver_timing.log, r1_jacobi_pdl_experiment.log); they were removed.
#pragma once
#include "ptx.cuh"

namespace nls {

constexpr int JB = 4;       // block width
constexpr int JP = 2 * JB;  // indices per pivot pair

// Pair `slot` of round `round` in a round-robin tournament over n (even) players.
__device__ __forceinline__ void rr_pair(int n, int round, int slot, int& p, int& q) {
  const int r = n - 1;
  if (slot == 0) {
    p = r;
    q = round;
  } else {
    p = (round + slot) % r;
    q = (round - slot + r) % r;
  }
}

// Global index of local index a (0..7) of the pair (bp, bq) of 4-wide blocks.
__device__ __forceinline__ int pair_index(int bp, int bq, int a) { return a < JB ? bp * JB + a : bq * JB + a - JB; }

// G (planar, row-major, pitch ld) = scale * A (interleaved complex, m x m) zero-padded to mp; V = I.
__global__ void jacobi_init_kernel(const double* __restrict__ A, int m, int mp, double scale, double* __restrict__ Gr,
                                   double* __restrict__ Gi, double* __restrict__ Vr, double* __restrict__ Vi) {
  const long long total = (long long)mp * mp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % mp), r = (int)(e / mp);
    double re = 0.0, im = 0.0;
    if (r < m && c < m) {
      re = scale * A[((long long)r * m + c) * 2];
      im = scale * A[((long long)r * m + c) * 2 + 1];
    }
    Gr[e] = re;
    Gi[e] = (r == c) ? 0.0 : im;
    Vr[e] = (r == c) ? 1.0 : 0.0;
    Vi[e] = 0.0;
  }
}

// sum of squares of a vector (single block, fixed order).
__global__ void sumsq_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n,
                             double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (long long e = threadIdx.x; e < n; e += blockDim.x) acc += a[e] * a[e] + b[e] * b[e];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    out[0] = t;
  }
}

// thr[0] = (eps ||G||_F)^2 : squared absolute rotation threshold, thr[1] = 0 (relative test off).
__global__ void jacobi_threshold_kernel(const double* __restrict__ fro2, double* __restrict__ thr) {
  const double eps = 2.220446049250313e-16;
  thr[0] = eps * eps * fro2[0];
  thr[1] = 0.0;
}

// One warp: J = eigenvectors of the 8 x 8 pivot block G[I, I] (I = blocks bp, bq) by cyclic two-sided
// Jacobi in shared memory (32 lanes = 8 rows x 4 disjoint rotations per inner round).
// sm: [4][8][9] doubles (Sr, Si, Jr, Ji).  Writes J to out[2][8][8]; returns whether anything rotated.
__device__ __forceinline__ bool pivot_solve(const double* __restrict__ Gr, const double* __restrict__ Gi, int ld,
                                            int bp, int bq, double thr_abs2, double thr_rel2, int max_inner,
                                            double (*sm)[8][9], double* __restrict__ out, int lane) {
  double(*Sr)[9] = sm[0];
  double(*Si)[9] = sm[1];
  double(*Jr)[9] = sm[2];
  double(*Ji)[9] = sm[3];
  for (int e = lane; e < 64; e += 32) {
    const int a = e >> 3, b = e & 7;
    const long long o = (long long)pair_index(bp, bq, a) * ld + pair_index(bp, bq, b);
    Sr[a][b] = __ldcg(Gr + o);
    Si[a][b] = __ldcg(Gi + o);
    Jr[a][b] = (a == b) ? 1.0 : 0.0;
    Ji[a][b] = 0.0;
  }
  __syncwarp();
  const int i = lane >> 2, slot = lane & 3;
  bool any_total = false;
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    bool any = false;
    for (int r = 0; r < 7; ++r) {
      int p, q;
      rr_pair(8, r, slot, p, q);
      const double a = Sr[p][p], b = Sr[q][q], zr = Sr[p][q], zi = Si[p][q];
      const double az2 = zr * zr + zi * zi;
      // Rotate unless |g_pq| is below the absolute noise floor eps ||G||_F (and, optionally, negligible
      // relative to its diagonal pair).  A purely relative test never terminates on the numerically
      // rank-deficient Gram matrices of this path: their null-space block is rounding noise.
      const bool rot = az2 > thr_abs2 && az2 > thr_rel2 * fabs(a * b);
      double c = 1.0, s = 0.0, er = 1.0, ei = 0.0;
      if (rot) {
        const double inv_az = rsqrt(az2);
        const double tau = 0.5 * (b - a) * inv_az;
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        c = rsqrt(1.0 + t * t);
        s = t * c;
        er = zr * inv_az;
        ei = zi * inv_az;
      }
      any |= rot;
      const double wr = s * er, wi = s * ei;  // w = s e
      __syncwarp();
      {  // columns p, q of S and J, row i:  x_p' = c x_p - conj(w) x_q ;  x_q' = w x_p + c x_q
        const double pr = Sr[i][p], pi = Si[i][p], qr = Sr[i][q], qi = Si[i][q];
        Sr[i][p] = c * pr - (wr * qr + wi * qi);
        Si[i][p] = c * pi - (wr * qi - wi * qr);
        Sr[i][q] = (wr * pr - wi * pi) + c * qr;
        Si[i][q] = (wr * pi + wi * pr) + c * qi;
        const double jpr = Jr[i][p], jpi = Ji[i][p], jqr = Jr[i][q], jqi = Ji[i][q];
        Jr[i][p] = c * jpr - (wr * jqr + wi * jqi);
        Ji[i][p] = c * jpi - (wr * jqi - wi * jqr);
        Jr[i][q] = (wr * jpr - wi * jpi) + c * jqr;
        Ji[i][q] = (wr * jpi + wi * jpr) + c * jqi;
      }
      __syncwarp();
      {  // rows p, q of S, column i:  x_p' = c x_p - w x_q ;  x_q' = conj(w) x_p + c x_q
        const double pr = Sr[p][i], pi = Si[p][i], qr = Sr[q][i], qi = Si[q][i];
        Sr[p][i] = c * pr - (wr * qr - wi * qi);
        Si[p][i] = c * pi - (wr * qi + wi * qr);
        Sr[q][i] = (wr * pr + wi * pi) + c * qr;
        Si[q][i] = (wr * pi - wi * pr) + c * qi;
      }
      __syncwarp();
      if (rot && i == 0) {  // the rotated pivot is diagonal by construction: remove the rounding residue
        Sr[p][q] = Sr[q][p] = 0.0;
        Si[p][q] = Si[q][p] = 0.0;
        Si[p][p] = Si[q][q] = 0.0;
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, any)) break;
    any_total = true;
  }
  for (int e = lane; e < 64; e += 32) {
    out[e] = Jr[e >> 3][e & 7];
    out[64 + e] = Ji[e >> 3][e & 7];
  }
  return any_total;
}

// Two-kernel variant, kernel 1: one warp per pivot pair.
// Jbuf: [pair][2][8][8] (Re, Im), flags[pair] = 1 if any rotation was applied, *active += #rotating pairs.
__global__ void __launch_bounds__(128) jacobi_pivot_kernel(const double* __restrict__ Gr, const double* __restrict__ Gi,
                                                           int ld, int nb, int round,
                                                           const double* __restrict__ thr /* [abs^2, rel^2] */,
                                                           int max_inner, double* __restrict__ Jbuf,
                                                           int* __restrict__ flags, int* __restrict__ active) {
  __shared__ double sm[4][4][8][9];  // per warp: Sr, Si, Jr, Ji
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + warp;
  if (pair >= nb / 2) return;
  int bp, bq;
  rr_pair(nb, round, pair, bp, bq);
  const bool any = pivot_solve(Gr, Gi, ld, bp, bq, thr[0], thr[1], max_inner, sm[warp], Jbuf + (long long)pair * 128, lane);
  if (lane == 0) {
    flags[pair] = any ? 1 : 0;
    if (any) atomicAdd(active, 1);
  }
}

// Row-major enumeration of the upper triangle of an np x np tile grid (row i holds np - i tiles).
__device__ __forceinline__ void upper_tile(long long task, int np, int& i, int& j) {
  const double disc = (2.0 * np + 1.0) * (2.0 * np + 1.0) - 8.0 * (double)task;
  i = (int)((2.0 * np + 1.0 - sqrt(disc)) * 0.5);
  long long first = (long long)i * np - (long long)i * (i - 1) / 2;
  while (first > task) {
    --i;
    first = (long long)i * np - (long long)i * (i - 1) / 2;
  }
  while (first + (np - i) <= task) {
    first += np - i;
    ++i;
  }
  j = i + (int)(task - first);
}

// One warp: tile (i, j) of G <- J^H G J (is_g; i <= j, also written conjugate-transposed to (j, i)) or row block
// i of V <- V J (!is_g), for the pairing of `round`, as complex 8x8x8 products on DMMA.8x8x4.  The tile is read
// and written by this warp only, so the update is in place.  ts: [2][8][9] doubles of per-warp scratch.
__device__ __forceinline__ void update_tile(double* __restrict__ Gr, double* __restrict__ Gi, double* __restrict__ Vr,
                                            double* __restrict__ Vi, int ld, int nb, int round,
                                            const double* __restrict__ Jbuf, const int* __restrict__ flags,
                                            bool is_g, int i, int j, double (*ts)[8][9], int lane,
                                            const int2* __restrict__ pair_tab = nullptr) {
  const bool fj = __ldcg(flags + j) != 0;
  const bool fi = is_g && __ldcg(flags + i) != 0;
  if (!fi && !fj) return;  // warp-uniform
  const int fr = lane >> 2, fk = lane & 3;
  int jp, jq, ip = 0, iq = 0;
  if (pair_tab) {  // blocks of every pair of this round, tabulated once per CTA (no integer divisions here)
    jp = pair_tab[j].x;
    jq = pair_tab[j].y;
    if (is_g) {
      ip = pair_tab[i].x;
      iq = pair_tab[i].y;
    }
  } else {
    rr_pair(nb, round, j, jp, jq);
    if (is_g) rr_pair(nb, round, i, ip, iq);
  }
  double* Mr = is_g ? Gr : Vr;
  double* Mi = is_g ? Gi : Vi;
  const int row = is_g ? pair_index(ip, iq, fr) : i * 8 + fr;
  // All global loads of the task are issued up front (one L2 round trip instead of two).
  const double* Jj = Jbuf + (long long)j * 128;
  const double* Ji_ = Jbuf + (long long)i * 128;
  double ar[2], ai[2], br[2], bi[2], cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const long long o = (long long)row * ld + pair_index(jp, jq, 4 * ks + fk);
    ar[ks] = __ldcg(Mr + o);
    ai[ks] = __ldcg(Mi + o);
    br[ks] = __ldcg(Jj + (4 * ks + fk) * 8 + fr);
    bi[ks] = __ldcg(Jj + 64 + (4 * ks + fk) * 8 + fr);
    if (is_g) {  // A operand of the second product: (J_i^H)[m = fr][k] = conj(J_i[k][fr])
      cr[ks] = __ldcg(Ji_ + (4 * ks + fk) * 8 + fr);
      ci[ks] = __ldcg(Ji_ + 64 + (4 * ks + fk) * 8 + fr);
    }
  }
  // ---- T = M_tile * J_j  (identity if pair j did not rotate) ----
  double tr[2] = {0.0, 0.0}, ti[2] = {0.0, 0.0};
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    dmma(tr[0], tr[1], ar[ks], br[ks]);
    dmma(tr[0], tr[1], -ai[ks], bi[ks]);
    dmma(ti[0], ti[1], ar[ks], bi[ks]);
    dmma(ti[0], ti[1], ai[ks], br[ks]);
  }
  double outr[2] = {tr[0], tr[1]}, outi[2] = {ti[0], ti[1]};
  if (is_g) {
    // ---- out = J_i^H * T ----
    double(*Tr)[9] = ts[0];
    double(*Ti)[9] = ts[1];
    __syncwarp();
    Tr[fr][2 * fk] = tr[0];
    Tr[fr][2 * fk + 1] = tr[1];
    Ti[fr][2 * fk] = ti[0];
    Ti[fr][2 * fk + 1] = ti[1];
    __syncwarp();
    outr[0] = outr[1] = outi[0] = outi[1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const double tbr = Tr[4 * ks + fk][fr], tbi = Ti[4 * ks + fk][fr];  // B operand: T[k][n = fr]
      dmma(outr[0], outr[1], cr[ks], tbr);
      dmma(outr[0], outr[1], ci[ks], tbi);
      dmma(outi[0], outi[1], cr[ks], tbi);
      dmma(outi[0], outi[1], -ci[ks], tbr);
    }
  }
  const int col = pair_index(jp, jq, 2 * fk);
  const long long o = (long long)row * ld + col;
  *reinterpret_cast<double2*>(Mr + o) = make_double2(outr[0], outr[1]);
  *reinterpret_cast<double2*>(Mi + o) = make_double2(outi[0], outi[1]);
  if (is_g && i != j) {  // mirror tile: G[c, r] = conj(G[r, c])
    Mr[(long long)col * ld + row] = outr[0];
    Mi[(long long)col * ld + row] = -outi[0];
    Mr[(long long)(col + 1) * ld + row] = outr[1];
    Mi[(long long)(col + 1) * ld + row] = -outi[1];
  }
}

// Two-kernel variant, kernel 2: G <- J^H G J (upper tiles + mirrors) and V <- V J, one warp per 8 x 8 tile.
__global__ void __launch_bounds__(256) jacobi_update_kernel(double* __restrict__ Gr, double* __restrict__ Gi,
                                                            double* __restrict__ Vr, double* __restrict__ Vi, int ld,
                                                            int nb, int round, const double* __restrict__ Jbuf,
                                                            const int* __restrict__ flags) {
  __shared__ double ts[8][2][8][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int np = nb / 2;
  const int rb = nb * JB / 8;  // 8-row blocks of V
  const long long n_g = (long long)np * (np + 1) / 2;
  const long long total = n_g + (long long)rb * np;
  for (long long task = (long long)blockIdx.x * 8 + warp; task < total; task += (long long)gridDim.x * 8) {
    const bool is_g = task < n_g;
    int i, j;
    if (is_g) {
      upper_tile(task, np, i, j);
    } else {
      i = (int)((task - n_g) / np);
      j = (int)((task - n_g) % np);
    }
    update_tile(Gr, Gi, Vr, Vi, ld, nb, round, Jbuf, flags, is_g, i, j, ts[warp], lane);
  }
}

// Round-robin bookkeeping: partner of block B in round rho, and the pair slot of block B in round r.
__device__ __forceinline__ int rr_partner(int n, int rho, int B) {
  const int R = n - 1;
  if (B == R) return rho;
  if (B == rho) return R;
  return ((2 * rho - B) % R + R) % R;
}
__device__ __forceinline__ int rr_slot(int n, int r, int B) {
  const int R = n - 1;
  if (B == R || B == r) return 0;
  const int s = ((B - r) % R + R) % R;
  return s <= n / 2 - 1 ? s : ((r - B) % R + R) % R;
}

struct JacobiArgs {
  double *Gr, *Gi, *Vr, *Vi;
  int ld, nb, max_inner, max_sweeps;
  const double* thr;     // [abs^2, rel^2]
  double* Jbuf;          // [2][np][128]
  int* flags;            // [2][np]
  int* active;           // [max_sweeps + 2] rotation counters per sweep window
};

// ---------------------------------------------------------------------------------------------
// Fused round (default): ONE kernel per round.  The CTAs first update the 2 np priority tiles that hold next
// round's pivot blocks and publish their completion on a counter; warp 0 of CTA p then waits for that
// counter and solves pivot p of round g + 1, while every other warp streams through the rest of update(g).
// The wait only ever depends on lower-numbered CTAs doing non-blocking work, so it cannot deadlock.
// Compared with the two-kernel round this removes one 4.4 us kernel boundary and hides the pivot latency.
//   state[0] = g (global round counter, advanced by the kernel), state[1 + (g & 3)] = priority tiles done.
// ---------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(256, MINB) jacobi_round_kernel(const JacobiArgs a, int* __restrict__ state) {
  __shared__ double ts[8][2][8][9];
  __shared__ double psm[4][8][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = a.nb, np = nb / 2, R = nb - 1, ld = a.ld;
  const int rb = nb * JB / 8;
  const long long n_g = (long long)np * (np + 1) / 2;
  const long long total = n_g + (long long)rb * np;
  const int g = __ldcg(state);  // written by the previous round's kernel
  const int r = g % R, rho = (r + 1) % R;
  const int cur = g & 1, nxt = cur ^ 1;
  const double* Jcur = a.Jbuf + (long long)cur * np * 128;
  const int* fcur = a.flags + cur * np;
  int* done = state + 1 + (g & 3);
  if (blockIdx.x == 0 && threadIdx.x == 0) state[1 + ((g + 2) & 3)] = 0;  // recycle a counter nobody is using
  // Round-robin tables of this round: blocks of every pair, and every block's partner in the next round.
  __shared__ int2 pair_tab[160];
  __shared__ int partner_tab[320];
  for (int e = threadIdx.x; e < np; e += 256) {
    int p, q;
    rr_pair(nb, r, e, p, q);
    pair_tab[e] = make_int2(p, q);
  }
  for (int e = threadIdx.x; e < nb; e += 256) partner_tab[e] = rr_partner(nb, rho, e);
  __syncthreads();
  // ---- priority tiles: all diagonal tiles + the tile joining the two blocks of each next-round pair
  const long long wid = (long long)blockIdx.x * 8 + warp;
  if (wid < 2 * np) {
    int i = -1, j = -1;
    if (wid < np) {
      i = j = (int)wid;
    } else {
      int bp, bq;
      rr_pair(nb, rho, (int)wid - np, bp, bq);
      const int sa = rr_slot(nb, r, bp), sb = rr_slot(nb, r, bq);
      if (sa != sb) {
        i = sa < sb ? sa : sb;
        j = sa < sb ? sb : sa;
        const int xi = sa < sb ? bp : bq;
        int ip, iq, jp, jq;
        rr_pair(nb, r, i, ip, iq);
        rr_pair(nb, r, j, jp, jq);
        if (xi != ip) {
          const int pa = rr_partner(nb, rho, ip);
          if (pa == jp || pa == jq) i = -1;  // the other next-round pair joining these two pairs owns the tile
        }
      }
    }
    if (i >= 0) update_tile(a.Gr, a.Gi, a.Vr, a.Vi, ld, nb, r, Jcur, fcur, true, i, j, ts[warp], lane, pair_tab);
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      atomicAdd(done, 1);
    }
  }
  const bool pivot_warp = warp == 0 && (int)blockIdx.x < np;
  if (pivot_warp) {
    // ---- pivot of round g + 1, as soon as every priority tile has landed
    if (lane == 0) {
      unsigned int spins = 0;
      while (true) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(done) : "memory");
        if (v >= 2 * np) break;
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();
      }
    }
    __syncwarp();
    int bp, bq;
    rr_pair(nb, rho, blockIdx.x, bp, bq);
    const bool any = pivot_solve(a.Gr, a.Gi, ld, bp, bq, a.thr[0], a.thr[1], a.max_inner, psm,
                                 a.Jbuf + ((long long)nxt * np + blockIdx.x) * 128, lane);
    if (lane == 0) {
      a.flags[nxt * np + blockIdx.x] = any ? 1 : 0;
      if (any) atomicAdd(a.active + (g + 1) / R, 1);
    }
  } else {
    // ---- the rest of update(g): the upper tiles are walked incrementally (no divisions), then V's tiles
    const long long n_workers = (long long)gridDim.x * 8 - np;
    const long long worker = wid - ((int)blockIdx.x < np ? blockIdx.x + 1 : np);
    long long task = worker;
    if (task < n_g) {
      int i, j;
      upper_tile(task, np, i, j);
      int pos = j - i;
      while (true) {
        if (i != j) {
          const int2 pi = pair_tab[i], pj = pair_tab[j];
          const int pa = partner_tab[pi.x], pb = partner_tab[pi.y];
          if (!(pa == pj.x || pa == pj.y || pb == pj.x || pb == pj.y))  // priority tiles were done above
            update_tile(a.Gr, a.Gi, a.Vr, a.Vi, ld, nb, r, Jcur, fcur, true, i, j, ts[warp], lane, pair_tab);
        }
        task += n_workers;
        if (task >= n_g) break;
        long long adv = n_workers + pos;
        while (adv >= np - i) {
          adv -= np - i;
          ++i;
        }
        pos = (int)adv;
        j = i + pos;
      }
    }
    {
      long long vt = task - n_g;  // first V task of this worker
      int i = (int)(vt / np), j = (int)(vt % np);
      const int di = (int)(n_workers / np), dj = (int)(n_workers % np);
      for (; vt < (long long)rb * np; vt += n_workers) {
        update_tile(a.Gr, a.Gi, a.Vr, a.Vi, ld, nb, r, Jcur, fcur, false, i, j, ts[warp], lane, pair_tab);
        i += di;
        j += dj;
        if (j >= np) {
          j -= np;
          ++i;
        }
      }
    }
  }
  // Advance the round counter once the whole grid is done with state[0]: the last CTA to finish does it.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int arrived = atomicAdd(state + 5, 1);
    if (arrived == (int)gridDim.x - 1) {
      state[5] = 0;
      __threadfence();
      state[0] = g + 1;
    }
  }
}

__global__ void jacobi_diag_kernel(const double* __restrict__ Gr, int ld, int m, double* __restrict__ lam) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < m) lam[k] = Gr[(long long)k * ld + k];
}

// Q_out[l, k] = V[l, perm[k]] (interleaved complex, m x m), lam_out[k] = lam_raw[perm[k]].
__global__ void jacobi_gather_kernel(const double* __restrict__ Vr, const double* __restrict__ Vi, int ld, int m,
                                     const int* __restrict__ perm, const double* __restrict__ lam_raw,
                                     double* __restrict__ Q_out, double* __restrict__ lam_out) {
  const long long total = (long long)m * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % m), l = (int)(e / m);
    const long long o = (long long)l * ld + perm[k];
    Q_out[2 * e] = Vr[o];
    Q_out[2 * e + 1] = Vi[o];
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) lam_out[k] = lam_raw[perm[k]];
}

}  // namespace nls
