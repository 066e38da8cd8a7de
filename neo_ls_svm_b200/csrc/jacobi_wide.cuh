// jacobi_wide.cuh — stage 3 eigensolver, default variant: parallel two-sided block Jacobi with 8-wide blocks
// (16 x 16 pivots), ping-pong G and ONE dependency-free kernel per round.
//
// Replaces scipy.linalg.eigh at _neo_ls_svm.py:120 for the m x m complex Hermitian matrix A/c.
//
// Why this shape (measured on B200, m = 1025):
//   * With 4-wide blocks (jacobi.cuh) a round moves all of G and V through L2 for an 8-deep product per element
//     and a sweep needs nb - 1 = 257 rounds: the sweep is bound by per-round latency, not by the FP64 pipe.
//     8-wide blocks halve the rounds and double the DMMA work per byte.
//   * The critical path of a round is: launch -> (tiles that hold the next pivots) -> pivot solve -> end.  Here
//     G is double-buffered (round g reads G[g & 1], writes G[(g + 1) & 1]), so nothing in a round depends on
//     anything else in the same round: the warp that solves next round's pivot (bp', bq') assembles its
//     16 x 16 block itself from the OLD G —
//         diagonal 8 x 8 blocks : sub-blocks of the final S of the current pivots (kept by their solver, Sbuf),
//         cross 8 x 8 block     : J_a[:, bp']^H  G_old[I_a, I_b]  J_b[:, bq']   (48 DMMAs),
//     — while every other warp streams through the tile updates.  No priority tiles, no spin-waits, no
//     co-residency requirement.
//   * The pivot solve is a chain of dependent FP64 operations.  It is cut three ways: one stage per inner round
//     (S is updated as (P/2)^2 independent 2 x 2 blocks R_a^H S_ab R_b, one __syncwarp), rotations from two
//     rsqrt (MUFU seed + Newton) instead of four library calls, and only the 64 CROSS pairs (i in bp, j in bq)
//     are annihilated per pivot — the pairs inside a block are visited in the first round of each sweep, which
//     is all a cyclic sweep asks for (every pair exactly once per sweep).
//
// Rotation threshold, convergence criterion and padding are those of jacobi.cuh: rotate iff |g_pq| > eps ||G||_F;
// stop when no off-diagonal entry is above it (census kernel; equivalent to a sweep that rotates nothing).
#pragma once
#include "jacobi.cuh"

namespace nls {

template <int JBW>
__device__ __forceinline__ int pair_index_w(int bp, int bq, int a) {
  return a < JBW ? bp * JBW + a : bq * JBW + a - JBW;
}

template <int JBW>
struct WideCfg {
  static constexpr int P = 2 * JBW;           // indices per pivot pair
  static constexpr int NS = P / 8;            // 8-wide DMMA sub-tiles per tile side
  static constexpr int KS = P / 4;            // k4 steps per product
  static constexpr int TP = P + 4;            // pitch of the per-warp T scratch (conflict-free B-fragment reads)
  static constexpr int SP = P + 1;            // pitch of the pivot scratch
  static constexpr int JSZ = 2 * P * P;       // doubles per J / per S (Re plane, Im plane; row-major)
  static constexpr int TS_WARP = 2 * P * TP;  // doubles of T scratch per warp
  static constexpr int PIV_SM = 4 * P * SP;   // doubles of pivot scratch (Sr, Si, Jr, Ji)
};

struct WideArgs {
  double* G[2][2];  // [buffer][Re, Im], ping-pong by round parity
  double *Vr, *Vi;
  int ld, nb, max_inner, sm_count;
  const double* thr;  // [abs^2]
  double* Jbuf;       // [2][np][JSZ]  J of the pivots, by round parity
  double* Sbuf;       // [2][np][JSZ]  final S of the pivot solves (= the updated diagonal tiles)
  int* flags;         // [2][np]       1 if the pair rotated
  int* active;        // [max_sweeps + 2] rotating pairs per sweep
  int* state;         // [0] sweeps completed (advanced by the census kernel that ends every sweep)
};

// 1/sqrt(x) for x in [2^-120, 2^120] (no exponent handling): MUFU.RSQ seed (2^-22) + ONE third-order step,
//   y = y0 (1 + r/2 + 3 r^2/8),  r = 1 - x y0^2  (truncation 5 r^3/16 < 1e-19), i.e. ~0.6 ulp.
__device__ __forceinline__ double rsqrt_seeded(double x) {
  const double y0 = (double)rsqrtf((float)x);
  const double r = fma(-x * y0, y0, 1.0);
  return fma(y0 * r, fma(0.375, r, 0.5), y0);
}

// Rotation that annihilates z = g_pq of the 2 x 2 Hermitian block [[a, z], [conj z, b]]:
//   x_p' = c x_p - conj(w) x_q ,  x_q' = w x_p + c x_q   (columns);   c = cos, w = sin * z / |z|.
// With d = (b - a)/2 and r = sqrt(d^2 + |z|^2):  (c, w) = (r + |d|, sign(d) z) / sqrt(2 r (r + |d|)) — the inner
// (|t| <= 1) rotation of the textbook tau/t formulas.  The library rsqrt()/sqrt()/division they need are four
// dependent ~40-instruction sequences; the pivot solve is the critical path of every round, so the rotation is
// computed from two seeded rsqrt in a power-of-two scaled domain (one exponent extraction).  Measured on B200
// (scripts/micro/rot_bias.cu): c^2 + |w|^2 - 1 has rms 1.5e-16 and mean < 1e-18, like the textbook form.
// Outside the scaled fast path's range (|d|^2 + |z|^2 not in [1e-280, 1e280]) the textbook form is used.
__device__ __forceinline__ void jacobi_rotation(double a, double b, double zr, double zi, double az2, double& c,
                                                double& wr, double& wi) {
  const double d = 0.5 * (b - a), ad = fabs(d);
  const double r2 = fma(d, d, az2);
  if (!(r2 > 1e-280 && r2 < 1e280)) {
    const double inv_az = rsqrt(az2);
    const double tau = d * inv_az;
    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
    c = rsqrt(1.0 + t * t);
    wr = t * c * zr * inv_az;
    wi = t * c * zi * inv_az;
    return;
  }
  const int e2 = (((__double2hiint(r2) >> 20) & 0x7ff) - 1023) & ~1;        // even exponent of r2
  const double s2 = __hiloint2double((1023 - e2) << 20, 0);                  // 2^-e2
  const double s1 = __hiloint2double((1023 - e2 / 2) << 20, 0);              // 2^-(e2/2)
  const double r2s = r2 * s2;                                                // in [1, 4)
  const double rs = r2s * rsqrt_seeded(r2s);                                 // r 2^-(e2/2) in [1, 2]
  const double rds = fma(ad, s1, rs);                                        // (r + |d|) 2^-(e2/2) in [1, 4]
  const double iq = rsqrt_seeded((rs + rs) * rds);                           // argument in [2, 32]
  const double sg = (d >= 0.0 ? iq : -iq) * s1;
  c = rds * iq;
  wr = sg * zr;
  wi = sg * zi;
}

// Rotation pairing of inner round r: full = round-robin over all P indices (P - 1 rounds); cross = index s of the
// first block with index (s + r) mod JBW of the second (JBW rounds).
template <int JBW>
__device__ __forceinline__ void inner_pair(bool cross, int r, int slot, int& p, int& q) {
  if (cross) {
    p = slot;
    q = JBW + ((slot + r) & (JBW - 1));
  } else {
    rr_pair(2 * JBW, r, slot, p, q);
  }
}

// Barrier among the NW warps that solve one pivot (named barrier 1: a CTA holds at most one multi-warp pivot).
template <int NW>
__device__ __forceinline__ void pivot_sync() {
  if (NW == 1)
    __syncwarp();
  else
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
}

// NW warps (tid = 0 .. 32 NW - 1): cyclic two-sided Jacobi on the P x P Hermitian block held in shared memory (Sr,
// Si at sm, pitch SP); accumulates J (Jr, Ji, initialised here).  An inner round applies P/2 disjoint rotations
// in ONE stage: S is updated as (P/2)^2 independent 2 x 2 blocks R_a^H S[{p_a,q_a},{p_b,q_b}] R_b (each thread owns
// blocks of its own column rotation b and fetches the row rotations by shuffle — every warp computes all P/2
// rotations), J as column pairs; one barrier between the reads and the writes, one after.  With NW = 2 the two
// warps sit on different SM sub-partitions and each does half of the FP64 and shared-memory work of a round.
// Returns whether anything rotated.
template <int JBW, int NW>
__device__ __forceinline__ bool pivot_rotate_w(double* __restrict__ sm, double thr_abs2, bool cross, int max_inner,
                                               int tid) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P, SP = C::SP, NROT = P / 2, T = 32 * NW;
  static_assert(NROT <= 32 && 32 % NROT == 0 && (JBW & (JBW - 1)) == 0, "every warp holds all rotations");
  constexpr int LPR = T / NROT;    // threads sharing a column rotation
  constexpr int BPL = NROT / LPR;  // 2 x 2 blocks of S per thread and round
  constexpr int RPL = P / LPR;     // rows of J per thread and round
  static_assert(BPL >= 1 && RPL >= 1, "too many warps for this pivot size");
  double* Sr = sm;
  double* Si = sm + P * SP;
  double* Jr = sm + 2 * P * SP;
  double* Ji = sm + 3 * P * SP;
  for (int e = tid; e < P * P; e += T) {
    const int a = e / P, b = e % P;
    Jr[a * SP + b] = (a == b) ? 1.0 : 0.0;
    Ji[a * SP + b] = 0.0;
  }
  pivot_sync<NW>();
  const int lane = tid;
  const int slot = lane % NROT, sub = lane / NROT;
  const int nrounds = cross ? JBW : P - 1;
  bool any_total = false;
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    bool any = false;
    for (int r = 0; r < nrounds; ++r) {
      int p, q;
      inner_pair<JBW>(cross, r, slot, p, q);
      const double zr = Sr[p * SP + q], zi = Si[p * SP + q], dp = Sr[p * SP + p], dq = Sr[q * SP + q];
      const double az2 = zr * zr + zi * zi;
      const bool rot = az2 > thr_abs2;  // absolute noise floor eps ||G||_F
      double c = 1.0, wr = 0.0, wi = 0.0;
      if (rot) jacobi_rotation(dp, dq, zr, zi, az2, c, wr, wi);
      any |= rot;
      // ---- S: blocks (a, slot) for the BPL row rotations a = sub + LPR * t
      double n00r[BPL], n00i[BPL], n01r[BPL], n01i[BPL], n10r[BPL], n10i[BPL], n11r[BPL], n11i[BPL];
      int pa[BPL], qa[BPL];
#pragma unroll
      for (int t = 0; t < BPL; ++t) {
        const int a = sub + LPR * t;
        const double ca = __shfl_sync(0xffffffffu, c, a), war = __shfl_sync(0xffffffffu, wr, a),
                     wai = __shfl_sync(0xffffffffu, wi, a);
        inner_pair<JBW>(cross, r, a, pa[t], qa[t]);
        const double x00r = Sr[pa[t] * SP + p], x00i = Si[pa[t] * SP + p], x01r = Sr[pa[t] * SP + q],
                     x01i = Si[pa[t] * SP + q], x10r = Sr[qa[t] * SP + p], x10i = Si[qa[t] * SP + p],
                     x11r = Sr[qa[t] * SP + q], x11i = Si[qa[t] * SP + q];
        // columns (rotation `slot`):  y_p = c x_p - conj(w) x_q ;  y_q = w x_p + c x_q
        const double y00r = c * x00r - (wr * x01r + wi * x01i), y00i = c * x00i - (wr * x01i - wi * x01r);
        const double y01r = (wr * x00r - wi * x00i) + c * x01r, y01i = (wr * x00i + wi * x00r) + c * x01i;
        const double y10r = c * x10r - (wr * x11r + wi * x11i), y10i = c * x10i - (wr * x11i - wi * x11r);
        const double y11r = (wr * x10r - wi * x10i) + c * x11r, y11i = (wr * x10i + wi * x10r) + c * x11i;
        // rows (rotation a):  z_p = c y_p - w y_q ;  z_q = conj(w) y_p + c y_q
        n00r[t] = ca * y00r - (war * y10r - wai * y10i);
        n00i[t] = ca * y00i - (war * y10i + wai * y10r);
        n10r[t] = (war * y00r + wai * y00i) + ca * y10r;
        n10i[t] = (war * y00i - wai * y00r) + ca * y10i;
        n01r[t] = ca * y01r - (war * y11r - wai * y11i);
        n01i[t] = ca * y01i - (war * y11i + wai * y11r);
        n11r[t] = (war * y01r + wai * y01i) + ca * y11r;
        n11i[t] = (war * y01i - wai * y01r) + ca * y11i;
        if (a == slot) {  // the pivot block itself: real diagonal, and annihilated by construction when rotated
          n00i[t] = n11i[t] = 0.0;
          if (rot) n01r[t] = n01i[t] = n10r[t] = n10i[t] = 0.0;
        }
      }
      // ---- J <- J R (rotation `slot`), rows sub + LPR * t
      double jp_r[RPL], jp_i[RPL], jq_r[RPL], jq_i[RPL];
#pragma unroll
      for (int t = 0; t < RPL; ++t) {
        const int i = sub + LPR * t;
        const double xr = Jr[i * SP + p], xi = Ji[i * SP + p], yr = Jr[i * SP + q], yi = Ji[i * SP + q];
        jp_r[t] = c * xr - (wr * yr + wi * yi);
        jp_i[t] = c * xi - (wr * yi - wi * yr);
        jq_r[t] = (wr * xr - wi * xi) + c * yr;
        jq_i[t] = (wr * xi + wi * xr) + c * yi;
      }
      pivot_sync<NW>();  // every thread has read its inputs
#pragma unroll
      for (int t = 0; t < BPL; ++t) {
        Sr[pa[t] * SP + p] = n00r[t];
        Si[pa[t] * SP + p] = n00i[t];
        Sr[pa[t] * SP + q] = n01r[t];
        Si[pa[t] * SP + q] = n01i[t];
        Sr[qa[t] * SP + p] = n10r[t];
        Si[qa[t] * SP + p] = n10i[t];
        Sr[qa[t] * SP + q] = n11r[t];
        Si[qa[t] * SP + q] = n11i[t];
      }
#pragma unroll
      for (int t = 0; t < RPL; ++t) {
        const int i = sub + LPR * t;
        Jr[i * SP + p] = jp_r[t];
        Ji[i * SP + p] = jp_i[t];
        Jr[i * SP + q] = jq_r[t];
        Ji[i * SP + q] = jq_i[t];
      }
      pivot_sync<NW>();
    }
    if (!__any_sync(0xffffffffu, any)) break;  // every warp sees all rotations: the decision is uniform
    any_total = true;
  }
  return any_total;
}

// Store the solver's J and final S (pitch SP in shared memory) as [2][P][P] each.
template <int JBW, int NW>
__device__ __forceinline__ void pivot_store_w(const double* __restrict__ sm, double* __restrict__ Jout,
                                              double* __restrict__ Sout, int tid) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P, SP = C::SP;
  for (int e = tid; e < P * P; e += 32 * NW) {
    const int o = (e / P) * SP + e % P;
    Sout[e] = sm[o];
    Sout[P * P + e] = sm[P * SP + o];
    Jout[e] = sm[2 * P * SP + o];
    Jout[P * P + e] = sm[3 * P * SP + o];
  }
}

// Pivots of the very first round (one warp per pair): S straight from G, full cyclic sweep.
template <int JBW>
__global__ void __launch_bounds__(128) jacobi_pivot0_w_kernel(const WideArgs a) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P, SP = C::SP;
  extern __shared__ __align__(16) double dyn_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 4 + warp, np = a.nb / 2;
  if (pair >= np) return;
  double* sm = dyn_sm + warp * C::PIV_SM;
  int bp, bq;
  rr_pair(a.nb, 0, pair, bp, bq);
  for (int e = lane; e < P * P; e += 32) {
    const int x = e / P, y = e % P;
    const long long o = (long long)pair_index_w<JBW>(bp, bq, x) * a.ld + pair_index_w<JBW>(bp, bq, y);
    sm[x * SP + y] = __ldcg(a.G[0][0] + o);
    sm[P * SP + x * SP + y] = __ldcg(a.G[0][1] + o);
  }
  __syncwarp();
  const bool any = pivot_rotate_w<JBW, 1>(sm, a.thr[0], false, a.max_inner, lane);
  pivot_store_w<JBW, 1>(sm, a.Jbuf + (long long)pair * C::JSZ, a.Sbuf + (long long)pair * C::JSZ, lane);
  if (lane == 0) {
    a.flags[pair] = any ? 1 : 0;
    if (any) atomicAdd(a.active, 1);
  }
}

// A unit of tile work of a round: upper tile (i, j) of G (is_g) or P-row block i x pair j of V.
struct TileTask {
  bool valid, is_g;
  int i, j;
};

__device__ __forceinline__ TileTask decode_tile_task(long long t, long long n_g, long long total, int np) {
  TileTask k;
  k.valid = t < total;
  k.is_g = t < n_g;
  k.i = k.j = 0;
  if (!k.valid) return k;
  if (k.is_g) {
    upper_tile(t, np, k.i, k.j);
  } else {
    k.i = (int)((t - n_g) / np);
    k.j = (int)((t - n_g) % np);
  }
  return k;
}

// Does the task read its source tile?  (Diagonal G tiles come from Sbuf; V tiles of a pair that did not rotate do
// not change.)
__device__ __forceinline__ bool tile_task_loads(const TileTask& k, const int* __restrict__ sflags) {
  return k.valid && (k.is_g ? k.i != k.j : sflags[k.j] != 0);
}

// One warp: start the asynchronous copy (cp.async, 16 B granules, L2 only) of the task's P x P source tile into a
// staging buffer [2][P][TP] (Re plane, Im plane).  One commit group per call, even when nothing is copied.
template <int JBW>
__device__ __forceinline__ void stage_tile_w(const TileTask& k, const double* Sr_, const double* Si_, int ld,
                                             const int2* __restrict__ pair_tab, const int* __restrict__ sflags,
                                             double* __restrict__ buf, int lane) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P, TP = C::TP, GPR = P / 2;  // granules (2 doubles) per tile row
  if (tile_task_loads(k, sflags)) {
    const int2 pj = pair_tab[k.j];
    const int2 pi = k.is_g ? pair_tab[k.i] : make_int2(0, 0);
#pragma unroll
    for (int e = lane; e < P * GPR; e += 32) {
      const int row = e / GPR, gq = e % GPR;
      const int grow = k.is_g ? pair_index_w<JBW>(pi.x, pi.y, row) : k.i * P + row;
      const long long o = (long long)grow * ld + pair_index_w<JBW>(pj.x, pj.y, 2 * gq);
      const uint32_t d = smem_u32(buf + row * TP + 2 * gq);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(Sr_ + o) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + P * TP * 8), "l"(Si_ + o) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// One warp: dst tile (i, j) = J_i^H M J_j, plus its conjugate-transposed mirror (is_g; i < j), or P-row block i of
// V <- V J_j (!is_g), where the source tile M has been staged in `buf` ([2][P][TP]): complex P x P x P products as
// (P/8)^2 DMMA.8x8x4 sub-tiles.  A G tile whose pairs did not rotate is copied.  `buf` doubles as the scratch that
// turns the first product into a B operand once the tile is in registers.
template <int JBW>
__device__ __forceinline__ void update_tile_w(double* Dr, double* Di, int ld, const double* __restrict__ Jbuf,
                                              const int* __restrict__ sflags, const TileTask& k,
                                              double* __restrict__ buf, int lane, const int2* __restrict__ pair_tab) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P, NS = C::NS, KS = C::KS, TP = C::TP;
  const bool is_g = k.is_g;
  const int i = k.i, j = k.j;
  const bool fj = sflags[j] != 0;
  const bool fi = is_g && sflags[i] != 0;
  const int fr = lane >> 2, fk = lane & 3;
  const int2 pj = pair_tab[j];
  const int2 pi = is_g ? pair_tab[i] : make_int2(0, 0);
  int rowi[NS];
#pragma unroll
  for (int rs = 0; rs < NS; ++rs) rowi[rs] = is_g ? pair_index_w<JBW>(pi.x, pi.y, rs * 8 + fr) : i * P + rs * 8 + fr;
  double ar[NS][KS], ai[NS][KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int rs = 0; rs < NS; ++rs) {
      ar[rs][ks] = buf[(rs * 8 + fr) * TP + 4 * ks + fk];
      ai[rs][ks] = buf[P * TP + (rs * 8 + fr) * TP + 4 * ks + fk];
    }
  if (!fi && !fj) {  // nothing rotated: carry the tile (and its mirror) over to the other buffer
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int col = pair_index_w<JBW>(pj.x, pj.y, 4 * ks + fk);
#pragma unroll
      for (int rs = 0; rs < NS; ++rs) {
        const long long o = (long long)rowi[rs] * ld + col, om = (long long)col * ld + rowi[rs];
        Dr[o] = ar[rs][ks];
        Di[o] = ai[rs][ks];
        Dr[om] = ar[rs][ks];
        Di[om] = -ai[rs][ks];
      }
    }
    return;
  }
  // ---- T = M_tile * J_j ----
  double tr[NS][NS][2], ti[NS][NS][2];
#pragma unroll
  for (int rs = 0; rs < NS; ++rs)
#pragma unroll
    for (int cs = 0; cs < NS; ++cs) tr[rs][cs][0] = tr[rs][cs][1] = ti[rs][cs][0] = ti[rs][cs][1] = 0.0;
  const double* Jj = Jbuf + (long long)j * C::JSZ;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int cs = 0; cs < NS; ++cs) {
      const double br = __ldg(Jj + (4 * ks + fk) * P + cs * 8 + fr);  // J is read-only here: L1 keeps it
      const double bi = __ldg(Jj + P * P + (4 * ks + fk) * P + cs * 8 + fr);
#pragma unroll
      for (int rs = 0; rs < NS; ++rs) {
        dmma(tr[rs][cs][0], tr[rs][cs][1], ar[rs][ks], br);
        dmma(tr[rs][cs][0], tr[rs][cs][1], -ai[rs][ks], bi);
        dmma(ti[rs][cs][0], ti[rs][cs][1], ar[rs][ks], bi);
        dmma(ti[rs][cs][0], ti[rs][cs][1], ai[rs][ks], br);
      }
    }
  }
  if (is_g) {
    // ---- out = J_i^H * T : T goes through shared memory to become a B operand ----
    double* Tr = buf;
    double* Ti = buf + P * TP;
    __syncwarp();  // every lane holds its part of M in registers: the staging buffer is free
#pragma unroll
    for (int rs = 0; rs < NS; ++rs)
#pragma unroll
      for (int cs = 0; cs < NS; ++cs) {
        const int o = (rs * 8 + fr) * TP + cs * 8 + 2 * fk;
        *reinterpret_cast<double2*>(Tr + o) = make_double2(tr[rs][cs][0], tr[rs][cs][1]);
        *reinterpret_cast<double2*>(Ti + o) = make_double2(ti[rs][cs][0], ti[rs][cs][1]);
        tr[rs][cs][0] = tr[rs][cs][1] = ti[rs][cs][0] = ti[rs][cs][1] = 0.0;
      }
    __syncwarp();
    const double* Ji_ = Jbuf + (long long)i * C::JSZ;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      double tbr[NS], tbi[NS];
#pragma unroll
      for (int cs = 0; cs < NS; ++cs) {  // B operand: T[k][n = cs*8 + fr]
        tbr[cs] = Tr[(4 * ks + fk) * TP + cs * 8 + fr];
        tbi[cs] = Ti[(4 * ks + fk) * TP + cs * 8 + fr];
      }
#pragma unroll
      for (int rs = 0; rs < NS; ++rs) {  // A operand: (J_i^H)[m = rs*8 + fr][k] = conj(J_i[k][m])
        const double cr = __ldg(Ji_ + (4 * ks + fk) * P + rs * 8 + fr);
        const double ci = __ldg(Ji_ + P * P + (4 * ks + fk) * P + rs * 8 + fr);
#pragma unroll
        for (int cs = 0; cs < NS; ++cs) {
          dmma(tr[rs][cs][0], tr[rs][cs][1], cr, tbr[cs]);
          dmma(tr[rs][cs][0], tr[rs][cs][1], ci, tbi[cs]);
          dmma(ti[rs][cs][0], ti[rs][cs][1], cr, tbi[cs]);
          dmma(ti[rs][cs][0], ti[rs][cs][1], -ci, tbr[cs]);
        }
      }
    }
  }
#pragma unroll
  for (int cs = 0; cs < NS; ++cs) {
    const int col = pair_index_w<JBW>(pj.x, pj.y, cs * 8 + 2 * fk);
#pragma unroll
    for (int rs = 0; rs < NS; ++rs) {
      const long long o = (long long)rowi[rs] * ld + col;
      *reinterpret_cast<double2*>(Dr + o) = make_double2(tr[rs][cs][0], tr[rs][cs][1]);
      *reinterpret_cast<double2*>(Di + o) = make_double2(ti[rs][cs][0], ti[rs][cs][1]);
      if (is_g) {  // mirror tile: G[c, r] = conj(G[r, c])
        const long long om = (long long)col * ld + rowi[rs];
        Dr[om] = tr[rs][cs][0];
        Di[om] = -ti[rs][cs][0];
        Dr[om + ld] = tr[rs][cs][1];
        Di[om + ld] = -ti[rs][cs][1];
      }
    }
  }
}

// One warp: diagonal tile i of the destination G = the final S of pivot i (J_i^H S J_i, as left by its solver).
template <int JBW>
__device__ __forceinline__ void copy_diag_tile_w(const double* __restrict__ S, double* __restrict__ Dr,
                                                 double* __restrict__ Di, int ld, int2 pi, int lane) {
  constexpr int P = 2 * JBW;
  for (int e = lane; e < P * P; e += 32) {
    const int x = e / P, y = e % P;
    const long long o = (long long)pair_index_w<JBW>(pi.x, pi.y, x) * ld + pair_index_w<JBW>(pi.x, pi.y, y);
    Dr[o] = __ldcg(S + e);
    Di[o] = __ldcg(S + P * P + e);
  }
}

// NW warps (warps 0 .. NW - 1 of the CTA): warp 0 assembles next round's pivot block (bp, bq) in shared memory from
// the OLD G, the current J's and the current pivots' final S (see the header); all NW warps then solve it.
template <int JBW, int NW>
__device__ __forceinline__ void next_pivot_w(const WideArgs& a, const double* __restrict__ Gr,
                                             const double* __restrict__ Gi, int r, int rho, int slot_next, int cur,
                                             int sweep_of_next, double* __restrict__ ts, double* __restrict__ psm,
                                             int tid) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P, SP = C::SP, TP = C::TP, KS = C::KS;
  const int nb = a.nb, np = nb / 2, ld = a.ld, nxt = cur ^ 1;
  const int lane = tid & 31;
  if (tid < 32) {
    int bp, bq;
    rr_pair(nb, rho, slot_next, bp, bq);
    const int sa = rr_slot(nb, r, bp), sb = rr_slot(nb, r, bq);
    int xa, ya, xb, yb;
    rr_pair(nb, r, sa, xa, ya);
    rr_pair(nb, r, sb, xb, yb);
    const int oa = xa == bp ? 0 : JBW, ob = xb == bq ? 0 : JBW;  // local offsets of bp in pair sa, of bq in pair sb
    const double* Sa = a.Sbuf + ((long long)cur * np + sa) * C::JSZ;
    const double* Sb = a.Sbuf + ((long long)cur * np + sb) * C::JSZ;
    double* Sr = psm;
    double* Si = psm + P * SP;
    // Same-pair entries (both diagonal blocks; everything when sa == sb, i.e. nb == 2): from the final S.
    for (int e = lane; e < P * P; e += 32) {
      const int x = e / P, y = e % P;
      const bool x1 = x >= JBW, y1 = y >= JBW;
      if (sa != sb && x1 != y1) continue;
      const double* S = x1 ? Sb : Sa;
      const int lx = (x1 ? ob : oa) + (x & (JBW - 1)), ly = (y1 ? ob : oa) + (y & (JBW - 1));
      Sr[x * SP + y] = __ldcg(S + lx * P + ly);
      Si[x * SP + y] = __ldcg(S + P * P + lx * P + ly);
    }
    if (sa != sb) {
      // Cross block X = J_a[:, oa..]^H  G[I_a, I_b]  J_b[:, ob..]  (JBW x JBW; JBW == 8: one DMMA tile).
      static_assert(JBW == 8, "cross-block assembly is written for 8-wide blocks");
      const int fr = lane >> 2, fk = lane & 3;
      const double* Ja = a.Jbuf + ((long long)cur * np + sa) * C::JSZ;
      const double* Jb = a.Jbuf + ((long long)cur * np + sb) * C::JSZ;
      double mr[2][KS], mi[2][KS], br[KS], bi[KS], cr[KS], ci[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int col = pair_index_w<JBW>(xb, yb, 4 * ks + fk);
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          const long long o = (long long)pair_index_w<JBW>(xa, ya, rs * 8 + fr) * ld + col;
          mr[rs][ks] = __ldcg(Gr + o);
          mi[rs][ks] = __ldcg(Gi + o);
        }
        br[ks] = __ldcg(Jb + (4 * ks + fk) * P + ob + fr);
        bi[ks] = __ldcg(Jb + P * P + (4 * ks + fk) * P + ob + fr);
        cr[ks] = __ldcg(Ja + (4 * ks + fk) * P + oa + fr);
        ci[ks] = __ldcg(Ja + P * P + (4 * ks + fk) * P + oa + fr);
      }
      double tr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, ti[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          dmma(tr[rs][0], tr[rs][1], mr[rs][ks], br[ks]);
          dmma(tr[rs][0], tr[rs][1], -mi[rs][ks], bi[ks]);
          dmma(ti[rs][0], ti[rs][1], mr[rs][ks], bi[ks]);
          dmma(ti[rs][0], ti[rs][1], mi[rs][ks], br[ks]);
        }
      double* Tr = ts;
      double* Ti = ts + P * TP;
      __syncwarp();
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        *reinterpret_cast<double2*>(Tr + (rs * 8 + fr) * TP + 2 * fk) = make_double2(tr[rs][0], tr[rs][1]);
        *reinterpret_cast<double2*>(Ti + (rs * 8 + fr) * TP + 2 * fk) = make_double2(ti[rs][0], ti[rs][1]);
      }
      __syncwarp();
      double xr[2] = {0.0, 0.0}, xi[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double tbr = Tr[(4 * ks + fk) * TP + fr], tbi = Ti[(4 * ks + fk) * TP + fr];
        dmma(xr[0], xr[1], cr[ks], tbr);
        dmma(xr[0], xr[1], ci[ks], tbi);
        dmma(xi[0], xi[1], cr[ks], tbi);
        dmma(xi[0], xi[1], -ci[ks], tbr);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // X[fr][2 fk + h] and its conjugate transpose
        Sr[fr * SP + JBW + 2 * fk + h] = xr[h];
        Si[fr * SP + JBW + 2 * fk + h] = xi[h];
        Sr[(JBW + 2 * fk + h) * SP + fr] = xr[h];
        Si[(JBW + 2 * fk + h) * SP + fr] = -xi[h];
      }
    }
  }
  pivot_sync<NW>();
  // The pairs inside a block are rotated in the first round of a sweep only; every other pivot annihilates its
  // 64 cross pairs.  (nb == 2: the single pivot is the whole matrix, always a full sweep.)
  const bool cross = rho != 0 && nb > 2;
  const bool any = pivot_rotate_w<JBW, NW>(psm, a.thr[0], cross, a.max_inner, tid);
  pivot_store_w<JBW, NW>(psm, a.Jbuf + ((long long)nxt * np + slot_next) * C::JSZ,
                         a.Sbuf + ((long long)nxt * np + slot_next) * C::JSZ, tid);
  if (tid == 0) {
    a.flags[nxt * np + slot_next] = any ? 1 : 0;
    if (any) atomicAdd(a.active + sweep_of_next, 1);
  }
}

// One round: G[nxt] = J^H G[cur] J (tile by tile), V <- V J in place, and the pivots of the next round.
// Warps 0 and 1 of CTA p < np (two SM sub-partitions) solve next round's pivot p together; every other warp strides
// over the tile list (upper tiles of G first: they cost twice a V tile).  Nothing in the kernel waits for anything
// else in it.  Layouts that keep the pivot warps away from the tile updates' DMMAs (dedicated pivot CTAs at 8 pivots
// per SM, or 16-warp CTAs with the pivot's SM sub-partition otherwise idle) were measured and were not faster at
// m = 1025: what the pivot chain gains, the tile updates lose in workers (profiles/r1_jacobi_wide_layouts.log).
// Every tile-update warp runs a two-stage pipeline over its tasks: the source tile of task t + 1 is on its way into
// shared memory (cp.async) while task t is in the DMMA pipe (ncu before this: long-scoreboard stalls 8.4 warps per
// issue against 3.8 for the math pipe, DMMA pipe 42 % active).
// Dynamic shared memory: 8 warps x 2 staging buffers (TS_WARP doubles each) + PIV_SM doubles.
constexpr int WIDE_WARPS = 8;
constexpr int WIDE_PIVOT_WARPS = 2;  // warps 0 and 1 of a pivot CTA (two SM sub-partitions) share one pivot solve
template <int JBW>
constexpr int wide_smem_doubles() {
  return WIDE_WARPS * 2 * WideCfg<JBW>::TS_WARP + WideCfg<JBW>::PIV_SM;
}

template <int JBW>
__global__ void __launch_bounds__(WIDE_WARPS * 32, 2) jacobi_round_w_kernel(const WideArgs a, const int r,
                                                                            const int diag) {
  using C = WideCfg<JBW>;
  constexpr int P = C::P;
  extern __shared__ __align__(16) double dyn_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = a.nb, np = nb / 2, R = nb - 1, ld = a.ld;
  const int rb = nb * JBW / P;  // P-row blocks of V
  const long long n_g = (long long)np * (np + 1) / 2;
  const int g = __ldcg(a.state) * R + r;  // global round number (R is odd: the buffer parity flips every sweep)
  const int rho = r + 1 == R ? 0 : r + 1;
  const int cur = g & 1;
  const double* Jcur = a.Jbuf + (long long)cur * np * C::JSZ;
  const int* fcur = a.flags + cur * np;
  const double *Sr_ = cur ? a.G[1][0] : a.G[0][0], *Si_ = cur ? a.G[1][1] : a.G[0][1];
  double *Dr = cur ? a.G[0][0] : a.G[1][0], *Di = cur ? a.G[0][1] : a.G[1][1];
  __shared__ int2 pair_tab[160];
  __shared__ int sflags[160];
  for (int e = threadIdx.x; e < np; e += WIDE_WARPS * 32) {
    int p, q;
    rr_pair(nb, r, e, p, q);
    pair_tab[e] = make_int2(p, q);
    sflags[e] = __ldcg(fcur + e);
  }
  __syncthreads();
  double* stage = dyn_sm + warp * 2 * C::TS_WARP;
  const bool pivot_cta = (int)blockIdx.x < np;
  if (pivot_cta && warp < WIDE_PIVOT_WARPS) {
    if (!(diag & 1))
      next_pivot_w<JBW, WIDE_PIVOT_WARPS>(a, Sr_, Si_, r, rho, blockIdx.x, cur, (g + 1) / R, dyn_sm,
                                          dyn_sm + WIDE_WARPS * 2 * C::TS_WARP, threadIdx.x);
    return;
  }
  if (diag & 2) return;
  // Tile-update warps are numbered over the grid, skipping the pivot warps and (experiment, diag bits 2 and 3)
  // the warps that would share the pivot warp's SM sub-partition: warp 4 of the pivot CTAs, and warps 0 and 4 of
  // the CTAs that are presumably co-resident with them (blockIdx.x + sm_count).
  constexpr unsigned PIV = (1u << WIDE_PIVOT_WARPS) - 1u;  // pivot warps of a pivot CTA
  const unsigned mask_p = 0xFFu & ~(PIV | ((diag & 4) ? PIV << 4 : 0u)), mask_q = (diag & 8) ? 0xFFu & ~(PIV | PIV << 4) : 0xFFu;
  const int ps = a.sm_count, gx = (int)gridDim.x, bx = (int)blockIdx.x;
  auto clampi = [](int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); };
  auto workers_before = [&](int b) {
    const int q_lo = clampi(ps, np, gx), q_hi = clampi(ps + np, np, gx);
    return (long long)__popc(mask_p) * clampi(b, 0, np) + 8LL * (clampi(b, np, q_lo) - np) +
           (long long)__popc(mask_q) * (clampi(b, q_lo, q_hi) - q_lo) + 8LL * (clampi(b, q_hi, gx) - q_hi);
  };
  const unsigned my_mask = pivot_cta ? mask_p : ((bx >= ps && bx < ps + np) ? mask_q : 0xFFu);
  if (!((my_mask >> warp) & 1u)) return;
  const long long n_workers = workers_before(gx);
  const long long total = n_g + (long long)rb * np;
  long long t = workers_before(bx) + __popc(my_mask & ((1u << warp) - 1u));
  TileTask k = decode_tile_task(t, n_g, total, np);
  int b = 0;
  stage_tile_w<JBW>(k, k.is_g ? Sr_ : a.Vr, k.is_g ? Si_ : a.Vi, ld, pair_tab, sflags, stage, lane);
  while (k.valid) {
    const TileTask kn = decode_tile_task(t + n_workers, n_g, total, np);
    stage_tile_w<JBW>(kn, kn.is_g ? Sr_ : a.Vr, kn.is_g ? Si_ : a.Vi, ld, pair_tab, sflags,
                      stage + (b ^ 1) * C::TS_WARP, lane);
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // the tile of task k has landed (per thread) ...
    __syncwarp();                                          // ... for every lane of the warp
    if (k.is_g && k.i == k.j)
      copy_diag_tile_w<JBW>(a.Sbuf + ((long long)cur * np + k.i) * C::JSZ, Dr, Di, ld, pair_tab[k.i], lane);
    else if (tile_task_loads(k, sflags))
      update_tile_w<JBW>(k.is_g ? Dr : a.Vr, k.is_g ? Di : a.Vi, ld, Jcur, sflags, k, stage + b * C::TS_WARP, lane,
                         pair_tab);
    __syncwarp();  // the buffer of task k is refilled two iterations from now, by the next call of stage_tile_w
    k = kn;
    t += n_workers;
    b ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Number of off-diagonal entries of the current G still above the rotation threshold: zero means that a further
// sweep would rotate nothing, i.e. the iteration has converged (replaces the final no-op sweep).
__global__ void jacobi_offdiag_count_w_kernel(const WideArgs a, int mp, int* __restrict__ count) {
  const int cur = (__ldcg(a.state) + 1) & 1;  // the sweep that just ended ran an odd number of rounds
  const double *Gr = cur ? a.G[1][0] : a.G[0][0], *Gi = cur ? a.G[1][1] : a.G[0][1];
  const long long total = (long long)mp * mp;
  const double thr2 = a.thr[0];
  int c = 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(e % mp), row = (int)(e / mp);
    if (row == col) continue;
    const double re = __ldcg(Gr + e), im = __ldcg(Gi + e);
    c += (re * re + im * im > thr2) ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// Ends a sweep: state[0] += 1 (its own launch, after the census, so that every reader sees one value per kernel).
__global__ void jacobi_sweep_done_w_kernel(int* __restrict__ state) { state[0] += 1; }

}  // namespace nls
