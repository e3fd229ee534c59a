// Batched one-sided (Hestenes) block-Jacobi SVD for the 2-qubit gate step, complex128, sm_100a.
//
// Replaces exatn::decomposeTensorSVDLRSync (LAPACK zgesvd/zgesdd inside TAL-SH) at
// ExaTnMpsVisitor.cpp:1619-1627 and the partial-norm / slice truncation of :2366-2536.
//
// One launch = one "step" of a round-robin tournament over 8-column blocks; each CTA owns a pair of
// blocks (16 columns) of one matrix of the batch:
//   phase A  W = X^H X (16x16 Hermitian Gram matrix) on the FP64 tensor cores (DMMA) streaming the 16
//            columns from L2 -- the A- and B-fragments of X^T X coincide, so each element is loaded once;
//   phase B  one cyclic sweep of 2x2 Hermitian Jacobi rotations on W in shared memory (warp 0,
//            8 disjoint pairs per round), accumulating the 16x16 unitary Q;
//   phase C  X <- X Q on DMMA, written back in place.
// A CTA whose fresh Gram matrix is already orthogonal to tolerance does nothing; a matrix is converged
// after a full tournament without any rotation.  V is never accumulated: the other factor comes from
// one GEMM against the saved theta (see engine.cu), which is what keeps the kernel at 16 columns/CTA.
#include "kernels.h"
#include "ptx.cuh"
#include <algorithm>
#include <cstdio>

namespace mpsb200 {
namespace {

constexpr int JT = 128;     // threads per CTA of the 4-warp pair task (the 8-warp variant launches 256)
constexpr int WLD = 17;     // padded leading dimension of the 16x16 shared matrices
__device__ unsigned long long g_phase_cycles[8];   // developer timing (g_dbg_mode == 10): A, reduce+test, B, C, wait, tasks
__device__ unsigned long long g_dmma_flops = 0;   // real flops executed on the DMMA pipe by the pair tasks (reporting only)
__device__ int g_dbg_mode = 0;   // developer switches: 4 = rotate the pairs inside a block at every step (A/B run); 10 = per-phase clock64 timing

// barriers over the first 64 threads of the CTA (warps 0-1), named barrier 1: the rotation phase does not stop the other warps
__device__ __forceinline__ int bar64_or(int pred) {
  int any;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %1, 0;\n\tbar.red.or.pred p, 1, 64, q;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(any) : "r"(pred) : "memory");
  return any;
}
__device__ __forceinline__ void bar64_sync() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 rmul(double r, double2 a) { return make_double2(r * a.x, r * a.y); }

// round-robin tournament: the pair of 8-column blocks that task `pi` of step `step` owns (-1 = phantom block)
__device__ __forceinline__ bool pair_blocks(const JacobiProblem& P, int step, int pi, int& blkA, int& blkB, bool& within) {
  blkA = 0; blkB = -1; within = true;
  if (P.nb > 1) {
    const int nm1 = P.nbe - 1;
    const int s = step % nm1;
    within = (s == 0);
    if (pi == 0) { blkA = nm1; blkB = s; }
    else { blkA = (s + pi) % nm1; blkB = (s + nm1 - pi) % nm1; }
    if (blkA >= P.nb) blkA = -1;
    if (blkB >= P.nb) blkB = -1;
    if (blkA < 0) { blkA = blkB; blkB = -1; }
    if (blkA < 0) return false;
  }
  return true;
}

// one pair task: Gram (phase A), rotations (phase B), apply (phase C) on the 16 columns of blocks (blkA, blkB).
// All exits are uniform over the CTA.  Loads of G bypass L1 (ld.global.cg): inside the persistent sweep kernel the
// columns were last written by a CTA on another SM.
// NWARP = warps per task: 4 when the SMs hold several tasks each, 8 (16) when a tournament step has at most two (one) tasks per SM (layers
// of one to four gates in routed circuits) and the latency of the two streaming phases of a lone task is what counts.
//
// Phase C applies the rotations of phase B to the rows of X directly (one thread per row, the 16 entries of the row in
// registers) instead of accumulating a dense 16x16 Q and multiplying by it: on B200 the FP64 tensor pipe has the same
// FMA rate as the scalar FP64 pipe (both ~37 TFLOP/s), so what counts is the FMA count, and 64 plane rotations in their
// scaled ("fast Givens") form cost 64 x 8 + 32 = 544 FMA per row against 1024 for the dense complex 16x16 product.
// Every column carries a real scale gamma (product of the cosines it has been through in this task): with x = gamma x',
//   x'_p <- x'_p - tp x'_q,  x'_q <- x'_q + tq x'_p(old),   tp = conj(s/c) gamma_q / gamma_p,  tq = (s/c) gamma_p / gamma_q,
// and x = gamma x' once at the end.  gamma stays within [2^-7.5, 1] (|angle| <= pi/4, at most 15 rotations per column).
constexpr int MAXROT = 120;   // 8 cross rounds (+ 7 in-block rounds at the first step of a tournament) x 8 rotations
template <int NWARP>
__device__ __forceinline__ void pair_task(const JacobiProblem& P, int mat, int blkA, int blkB, bool within, double tol2, double dead2,
                                          const double* __restrict__ fro2, int* __restrict__ dirty) {
  const int M = P.M, N = P.N, ldg = P.ldg;
  double2* G = P.G;
  // columns whose squared norm is below dead_abs = (null_tol * ||G||_F)^2 are numerically null: they are zeroed at write-back
  // (trunc_kernel applies the same threshold to the singular values) and never rotated, so rank-deficient thetas do not spend
  // sweeps orthogonalising rounding noise.  The threshold is relative to the Frobenius norm, not to ||G||_F / sqrt(N): the
  // null columns of a rank-deficient 1024-column theta sit at ~1e-13 ||G||_F after a few sweeps (eps x sigma_max x the rotations
  // that touched them), and with the lower threshold they kept rotating against each other for ever (profiles/r04n_*)
  const double dead_abs = dead2 * fro2[mat];   // (null_tol * ||G||_F)^2: above the rounding-noise floor eps * sigma_max * sqrt(rotations) of a null column

  __shared__ int s_cols[16];
  constexpr int NT = 32 * NWARP;
  constexpr int NRED = NWARP < 8 ? NWARP : 8;   // partial-sum slots: sixteen-warp tasks fold warps 8-15 onto 0-7 (static shared memory stays < 48 KB)
  __shared__ double s_red[NRED][7][64];
  __shared__ double2 sW[16 * WLD];
  __shared__ double2 s_tp[MAXROT], s_tq[MAXROT];   // scaled rotation parameters in application order
  __shared__ double s_gam[16], s_igam[16];
  __shared__ double s_rc[8];
  __shared__ double2 s_rs[8];
  __shared__ int s_rp[8], s_rq[8];
  __shared__ int s_need;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 16) {
    int c;
    if (tid < 8) c = blkA * 8 + tid;
    else c = (blkB >= 0) ? blkB * 8 + (tid - 8) : N;
    s_cols[tid] = (c < N) ? c : -1;
    s_gam[tid] = 1.0; s_igam[tid] = 1.0;
  }
  if (tid == 0) s_need = 0;
  __syncthreads();
  if (!within && blkB < 0) return;   // a lone block has no cross pairs
  // Clean-pair memo: ver[b] counts the rotations block b has been through; rec[A][B] remembers the two versions at which the
  // cross pairs of (A, B) were last found orthogonal.  If neither block has changed since, the task is over before it loads
  // a single column -- this is what makes the last sweeps of a nearly converged matrix (and the final confirming sweep) cheap.
  int verA = 0, verB = 0;
  int2* recp = nullptr;
  if (!within) {
    verA = __ldcg(P.ver + blkA); verB = __ldcg(P.ver + blkB);
    recp = P.rec + (size_t)min(blkA, blkB) * P.nbe + max(blkA, blkB);
    const int2 rec = __ldcg(recp);
    const int va = blkA < blkB ? verA : verB, vb = blkA < blkB ? verB : verA;
    if (rec.x == va + 1 && rec.y == vb + 1) return;   // uniform: every thread read the same words
  }
  const bool timing = (g_dbg_mode == 10) && tid == 0;
  long long tA = 0, tB = 0, tC = 0, tD = 0;
  if (timing) tA = clock64();

  // ------------------------------------------------------------------ phase A: Gram matrix on DMMA
  {
    const int slot = lane >> 2, rsub = lane & 3;
    const int c0 = s_cols[slot], c1 = s_cols[8 + slot];
    const double2* p0 = (c0 >= 0) ? G + (size_t)ldg * c0 : nullptr;
    const double2* p1 = (c1 >= 0) ? G + (size_t)ldg * c1 : nullptr;
    double w00[2] = {0, 0}, m00[2] = {0, 0}, w11[2] = {0, 0}, m11[2] = {0, 0}, w01[2] = {0, 0}, p01[2] = {0, 0}, q01[2] = {0, 0};
    const int nch = (M + 3) >> 2;
    constexpr int UN = 4;
    if (!within) {
      // only the cross block W_AB = X_A^H X_B is computed (4 DMMA per 4 rows instead of 10): the Gram blocks W_AA, W_BB
      // of the two column blocks travel with them (P.wd), kept up to date by the two-sided rotations of every task and
      // recomputed from the columns at the first step of each tournament
      for (int base = warp * UN; base < nch; base += NWARP * UN) {
        double2 x0[UN], x1[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int row = 4 * (base + u) + rsub;
          const bool ok = row < M;
          x0[u] = (ok && p0) ? __ldcg(p0 + row) : make_double2(0.0, 0.0);
          x1[u] = (ok && p1) ? __ldcg(p1 + row) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          dmma884(w01[0], w01[1], x0[u].x, x1[u].x);
          dmma884(p01[0], p01[1], x0[u].x, x1[u].y);
          dmma884(q01[0], q01[1], x0[u].y, x1[u].x);
          dmma884(w01[0], w01[1], x0[u].y, x1[u].y);
        }
      }
    } else
    for (int base = warp * UN; base < nch; base += NWARP * UN) {
      double2 x0[UN], x1[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int row = 4 * (base + u) + rsub;
        const bool ok = row < M;
        x0[u] = (ok && p0) ? __ldcg(p0 + row) : make_double2(0.0, 0.0);
        x1[u] = (ok && p1) ? __ldcg(p1 + row) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        dmma884(w00[0], w00[1], x0[u].x, x0[u].x);
        dmma884(m00[0], m00[1], x0[u].x, x0[u].y);
        dmma884(w11[0], w11[1], x1[u].x, x1[u].x);
        dmma884(m11[0], m11[1], x1[u].x, x1[u].y);
        dmma884(w01[0], w01[1], x0[u].x, x1[u].x);
        dmma884(p01[0], p01[1], x0[u].x, x1[u].y);
        dmma884(q01[0], q01[1], x0[u].y, x1[u].x);
        dmma884(w00[0], w00[1], x0[u].y, x0[u].y);
        dmma884(w11[0], w11[1], x1[u].y, x1[u].y);
        dmma884(w01[0], w01[1], x0[u].y, x1[u].y);
      }
    }
    // C fragment: element (row = lane>>2, col = 2*(lane&3)+e)
    const int e0 = (lane >> 2) * 8 + 2 * (lane & 3);
    if (warp < NRED) {
      s_red[warp][0][e0] = w00[0]; s_red[warp][0][e0 + 1] = w00[1];
      s_red[warp][1][e0] = m00[0]; s_red[warp][1][e0 + 1] = m00[1];
      s_red[warp][2][e0] = w11[0]; s_red[warp][2][e0 + 1] = w11[1];
      s_red[warp][3][e0] = m11[0]; s_red[warp][3][e0 + 1] = m11[1];
      s_red[warp][4][e0] = w01[0]; s_red[warp][4][e0 + 1] = w01[1];
      s_red[warp][5][e0] = p01[0]; s_red[warp][5][e0 + 1] = p01[1];
      s_red[warp][6][e0] = q01[0]; s_red[warp][6][e0 + 1] = q01[1];
    }
    if (NWARP > NRED) {   // second half of the warps adds onto the first half's slots (fixed order: deterministic)
      __syncthreads();
      if (warp >= NRED) {
        const int w2 = warp - NRED;
        s_red[w2][0][e0] += w00[0]; s_red[w2][0][e0 + 1] += w00[1];
        s_red[w2][1][e0] += m00[0]; s_red[w2][1][e0 + 1] += m00[1];
        s_red[w2][2][e0] += w11[0]; s_red[w2][2][e0 + 1] += w11[1];
        s_red[w2][3][e0] += m11[0]; s_red[w2][3][e0 + 1] += m11[1];
        s_red[w2][4][e0] += w01[0]; s_red[w2][4][e0 + 1] += w01[1];
        s_red[w2][5][e0] += p01[0]; s_red[w2][5][e0 + 1] += p01[1];
        s_red[w2][6][e0] += q01[0]; s_red[w2][6][e0 + 1] += q01[1];
      }
    }
  }
  __syncthreads();
  if (timing) tB = clock64();
  for (int i = tid + (within ? 0 : 4 * 64); i < 7 * 64; i += NT) {
    const int t = i >> 6, e = i & 63;
    double acc = s_red[0][t][e];
#pragma unroll
    for (int w = 1; w < NRED; ++w) acc += s_red[w][t][e];
    s_red[0][t][e] = acc;
  }
  __syncthreads();
  const double2* wdA = P.wd + (size_t)blkA * 64;
  const double2* wdB = P.wd + (size_t)(blkB >= 0 ? blkB : blkA) * 64;
  for (int i = tid; i < 256; i += NT) {
    const int p = i >> 4, q = i & 15;
    const int bp = p >> 3, bq = q >> 3, r = p & 7, c = q & 7;
    double re, im;
    if (bp == bq && !within) { const double2 v = __ldcg((bp ? wdB : wdA) + r * 8 + c); re = v.x; im = v.y; }
    else if (bp == 0 && bq == 0) { re = s_red[0][0][r * 8 + c]; im = s_red[0][1][r * 8 + c] - s_red[0][1][c * 8 + r]; }
    else if (bp == 1 && bq == 1) { re = s_red[0][2][r * 8 + c]; im = s_red[0][3][r * 8 + c] - s_red[0][3][c * 8 + r]; }
    else if (bp == 0) { re = s_red[0][4][r * 8 + c]; im = s_red[0][5][r * 8 + c] - s_red[0][6][r * 8 + c]; }
    else { re = s_red[0][4][c * 8 + r]; im = -(s_red[0][5][c * 8 + r] - s_red[0][6][c * 8 + r]); }
    sW[p * WLD + q] = make_double2(re, im);
  }
  __syncthreads();
  // fresh-Gram convergence test over the pairs this task is responsible for
  {
    int need = 0;
    for (int i = tid; i < 256; i += NT) {
      const int p = i >> 4, q = i & 15;
      if (p < q && (within || (p < 8 && q >= 8))) {   // pairs inside a block belong to the first step of the tournament
        const double a = sW[p * WLD + p].x, b = sW[q * WLD + q].x;
        const double2 g = sW[p * WLD + q];
        if (a > dead_abs && b > dead_abs && (g.x * g.x + g.y * g.y) > tol2 * a * b) need = 1;
      }
    }
    if (need) s_need = 1;   // benign race: all writers store 1
  }
  __syncthreads();
  if (timing) { tC = clock64(); atomicAdd(&g_phase_cycles[0], (unsigned long long)(tB - tA)); atomicAdd(&g_phase_cycles[1], (unsigned long long)(tC - tB)); atomicAdd(&g_phase_cycles[5], 1ull); }
  if (!s_need) {
    if (!within && tid == 0) {
      const int va = blkA < blkB ? verA : verB, vb = blkA < blkB ? verB : verA;
      *recp = make_int2(va + 1, vb + 1);
    }
    if (within && tid < 128) {   // the tournament's first step refreshes the travelling Gram blocks from the columns
      const int bb = tid >> 6, r = (tid >> 3) & 7, c = tid & 7;
      if (bb == 0 || blkB >= 0) P.wd[(size_t)(bb ? blkB : blkA) * 64 + r * 8 + c] = sW[(8 * bb + r) * WLD + 8 * bb + c];
    }
    if (tid == 0) atomicAdd(&g_dmma_flops, (unsigned long long)((M + 3) >> 2) * (within ? 5120ull : 2048ull));   // Gram: 10 (4) DMMA per 4 rows
    return;
  }
  const int nrounds = within ? 15 : 8;
  if (tid == 0) atomicAdd(&g_dmma_flops, (unsigned long long)((M + 3) >> 2) * (within ? 5120ull : 2048ull) + (unsigned long long)M * (unsigned long long)(2 * (nrounds * 64 + 32)));
  if (tid == 0) dirty[mat] = 1;

  // ------------------------------------------------------------------ phase B: Jacobi rotations on W (warps 0-1; the other warps only take the barriers)
  // Round r rotates 8 disjoint column pairs.  Rounds 0-7 are the bipartite schedule over the cross pairs (i, 8 + (i+r)%8);
  // rounds 8-14 (only when `within`: the first step of a tournament) are the two 8-column round-robins of the pairs
  // inside each block, which every later step of the sweep leaves alone.  Lanes 0-7 of warp 0 compute the rotations
  // (two rsqrt, no division or sqrt) and their scaled form for phase C, warps 0-1 apply them to W two-sidedly.
  {
    if (warp < 2)
    for (int r = 0; r < nrounds; ++r) {
      int rot = 0;
      if (tid < 8) {
        int p, q;
        if (r < 8) { p = tid; q = 8 + ((tid + r) & 7); }
        else {
          const int w = r - 8, l = tid & 3, off = (tid >> 2) * 8;
          p = (l == 0) ? 7 : (w + l) % 7;
          q = (w + 7 - l) % 7;
          if (p > q) { const int t = p; p = q; q = t; }
          p += off; q += off;
        }
        const double a = sW[p * WLD + p].x, b = sW[q * WLD + q].x;
        const double2 g = sW[p * WLD + q];
        const double g2 = g.x * g.x + g.y * g.y;
        double c = 1.0;
        double2 sg = make_double2(0.0, 0.0);
        double2 tp = make_double2(0.0, 0.0), tq = make_double2(0.0, 0.0);
        if (a > dead_abs && b > dead_abs && g2 > tol2 * a * b) {
          rot = 1;
          // cos 2t = |d|/h, sin 2t = 2|g|/h  (|t| <= pi/4):  c = sqrt((1 + |d|/h)/2),  s = sign(d) g / (h c)
          const double d = b - a;
          const double rh = rsqrt(d * d + 4.0 * g2);
          const double x = 0.5 * (1.0 + fabs(d) * rh);
          const double rx = rsqrt(x);   // 1 / c
          c = x * rx;
          sg = rmul(d >= 0.0 ? rh * rx : -(rh * rx), g);
          // scaled form: t = s / c; the columns carry gamma_p, gamma_q
          const double2 t = rmul(rx, sg);
          const double gp = s_gam[p], gq = s_gam[q], igp = s_igam[p], igq = s_igam[q];
          tp = rmul(gq * igp, make_double2(t.x, -t.y));
          tq = rmul(gp * igq, t);
          s_gam[p] = gp * c; s_gam[q] = gq * c;
          s_igam[p] = igp * rx; s_igam[q] = igq * rx;
        }
        s_rc[tid] = c; s_rs[tid] = sg; s_rp[tid] = p; s_rq[tid] = q;
        s_tp[r * 8 + tid] = tp; s_tq[r * 8 + tid] = tq;
      }
      if (!bar64_or(rot)) continue;   // named barrier over warps 0-1 only (the other warps wait once, below)
      if (warp < 2) {
        // W <- J^H W J   (J_a = [[c, s],[-conj(s), c]] on columns (p,q) of pair a); one 2x2 block pair per thread
        const int ia = tid >> 3, ib = tid & 7;
        const int pa = s_rp[ia], qa = s_rq[ia], pb = s_rp[ib], qb = s_rq[ib];
        const double ca = s_rc[ia], cb = s_rc[ib];
        const double2 sa = s_rs[ia], sb = s_rs[ib];
        const double2 w00 = sW[pa * WLD + pb], w01 = sW[pa * WLD + qb], w10 = sW[qa * WLD + pb], w11 = sW[qa * WLD + qb];
        const double2 t00 = csub(rmul(ca, w00), cmul(sa, w10));
        const double2 t01 = csub(rmul(ca, w01), cmul(sa, w11));
        const double2 t10 = cadd(cmulc(sa, w00), rmul(ca, w10));
        const double2 t11 = cadd(cmulc(sa, w01), rmul(ca, w11));
        // (t * J_b): col p = t[:,0]*cb - t[:,1]*conj(sb) ; col q = t[:,0]*sb + t[:,1]*cb
        sW[pa * WLD + pb] = csub(rmul(cb, t00), cmulc(sb, t01));
        sW[pa * WLD + qb] = cadd(cmul(sb, t00), rmul(cb, t01));
        sW[qa * WLD + pb] = csub(rmul(cb, t10), cmulc(sb, t11));
        sW[qa * WLD + qb] = cadd(cmul(sb, t10), rmul(cb, t11));
      }
      bar64_sync();
    }
  }
  __syncthreads();   // rotation parameters and the rotated W are complete
  if (tid == 0) {   // this task is the only owner of both blocks during this step
    if (within) { verA = __ldcg(P.ver + blkA); if (blkB >= 0) verB = __ldcg(P.ver + blkB); }
    P.ver[blkA] = verA + 1;
    if (blkB >= 0) P.ver[blkB] = verB + 1;
  }
  if (tid < 128) {
    const int bb = tid >> 6, r = (tid >> 3) & 7, c = tid & 7;
    if (bb == 0 || blkB >= 0) P.wd[(size_t)(bb ? blkB : blkA) * 64 + r * 8 + c] = sW[(8 * bb + r) * WLD + 8 * bb + c];
  }
  __syncthreads();
  if (timing) { tD = clock64(); atomicAdd(&g_phase_cycles[2], (unsigned long long)(tD - tC)); }

  // ------------------------------------------------------------------ phase C: the rotations applied to the rows of X, in place
  {
    // the two blocks are 8 consecutive columns each; nA / nB of them exist (N need not be a multiple of 8)
    double2* const baseA = G + (size_t)ldg * (blkA * 8);
    double2* const baseB = (blkB >= 0) ? G + (size_t)ldg * (blkB * 8) : G;
    const int nA = min(8, N - blkA * 8), nB = (blkB >= 0) ? min(8, N - blkB * 8) : 0;
    for (int row = tid; row < M; row += NT) {
      asm volatile("" ::: "memory");   // keeps the loop-invariant loads of s_tp / s_tq at their point of use (else: hoisted into local memory)
      double2 x[16];
      {
        const double2* pa = baseA + row;
        const double2* pb = baseB + row;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          x[j] = (j < nA) ? __ldcg(pa) : make_double2(0.0, 0.0);
          x[8 + j] = (j < nB) ? __ldcg(pb) : make_double2(0.0, 0.0);
          pa += ldg; pb += ldg;
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int p = i, q = 8 + ((i + r) & 7);
          const double2 tp = s_tp[r * 8 + i], tq = s_tq[r * 8 + i];
          const double2 xp = x[p], xq = x[q];
          x[p] = make_double2(fma(-tp.x, xq.x, fma(tp.y, xq.y, xp.x)), fma(-tp.x, xq.y, fma(-tp.y, xq.x, xp.y)));
          x[q] = make_double2(fma(tq.x, xp.x, fma(-tq.y, xp.y, xq.x)), fma(tq.x, xp.y, fma(tq.y, xp.x, xq.y)));
        }
      }
      if (within) {
#pragma unroll
        for (int w = 0; w < 7; ++w) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int l = i & 3, off = (i >> 2) * 8;
            int p = (l == 0) ? 7 : (w + l) % 7;
            int q = (w + 7 - l) % 7;
            if (p > q) { const int t = p; p = q; q = t; }
            p += off; q += off;
            const double2 tp = s_tp[(8 + w) * 8 + i], tq = s_tq[(8 + w) * 8 + i];
            const double2 xp = x[p], xq = x[q];
            x[p] = make_double2(fma(-tp.x, xq.x, fma(tp.y, xq.y, xp.x)), fma(-tp.x, xq.y, fma(-tp.y, xq.x, xp.y)));
            x[q] = make_double2(fma(tq.x, xp.x, fma(-tq.y, xp.y, xq.x)), fma(tq.x, xp.y, fma(tq.y, xp.x, xq.y)));
          }
        }
      }
      {
        double2* pa = baseA + row;
        double2* pb = baseB + row;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double ga = s_gam[j], gb = s_gam[8 + j];
          if (j < nA) *pa = make_double2(ga * x[j].x, ga * x[j].y);
          if (j < nB) *pb = make_double2(gb * x[8 + j].x, gb * x[8 + j].y);
          pa += ldg; pb += ldg;
        }
      }
    }
  }
  if (timing) { atomicAdd(&g_phase_cycles[3], (unsigned long long)(clock64() - tD)); atomicAdd(&g_phase_cycles[6], 1ull); }
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// A whole sweep (nsteps tournament steps) of the whole batch in ONE launch: persistent CTAs pull pair tasks from a global
// counter in (step, matrix, pair) order and wait on per-block progress flags instead of on a kernel boundary:
//   task (s, m, A, B) may start when progress[m][A] >= base + s and progress[m][B] >= base + s, and publishes base + s + 1.
// Tasks are dequeued in dependency order and a waiting CTA only waits on tasks dequeued before its own, which are
// finished or running on resident CTAs, so the scheme cannot deadlock.  Compared with one launch per step this removes
// ~60 launch boundaries per sweep and lets the Gram / rotate / apply phases of different pairs overlap on an SM.
template <int NWARP>
__global__ void __launch_bounds__(32 * NWARP, NWARP >= 16 ? 1 : 16 / NWARP) jacobi_sweep_kernel(const JacobiProblem* __restrict__ probs, int batch, int max_pairs, int nsteps,
                                                             int base, double tol2, double dead2, const double* __restrict__ fro2,
                                                             int* __restrict__ dirty, const int* __restrict__ done,
                                                             int* __restrict__ progress, int progress_stride, int* __restrict__ counter,
                                                             int* __restrict__ fault, const int* __restrict__ active) {
  __shared__ int s_task;
  // active (optional): [0] = number of matrices still rotating, [1..] their indices (written by jacobi_check_kernel)
  if (active) batch = __ldcg(active);
  const int per_step = batch * max_pairs;
  const int total = nsteps * per_step;
  for (;;) {
    if (threadIdx.x == 0) s_task = atomicAdd(counter, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();   // s_task may be overwritten from here on
    if (t >= total) return;
    const int step = t / per_step, r = t - step * per_step;
    const int slot = r / max_pairs, pi = r - slot * max_pairs;
    const int mat = active ? __ldcg(active + 1 + slot) : slot;
    if (done[mat]) continue;
    const JacobiProblem P = probs[mat];
    const int npairs = (P.nb == 1) ? 1 : P.nbe / 2;
    if (pi >= npairs) continue;
    int blkA, blkB;
    bool within;
    if (!pair_blocks(P, step, pi, blkA, blkB, within)) continue;
    int* prog = progress + (size_t)mat * progress_stride;
    long long tw = 0;
    if (threadIdx.x == 0 && g_dbg_mode == 10) tw = clock64();
    if (threadIdx.x == 0) {
      // bounded wait (~1 s): a scheduling bug must surface as an error on the host, never as a hung GPU
      const int need = base + step;
      unsigned spins = 0;
      while (ld_acquire(prog + blkA) < need && ++spins < (1u << 23)) __nanosleep(64);
      if (blkB >= 0)
        while (ld_acquire(prog + blkB) < need && ++spins < (1u << 23)) __nanosleep(64);
      if (spins >= (1u << 23)) atomicAdd(fault, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && g_dbg_mode == 10) atomicAdd(&g_phase_cycles[4], (unsigned long long)(clock64() - tw));
    pair_task<NWARP>(P, mat, blkA, blkB, within || g_dbg_mode == 4, tol2, dead2, fro2, dirty);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      st_release(prog + blkA, base + step + 1);
      if (blkB >= 0) st_release(prog + blkB, base + step + 1);
    }
  }
}

__global__ void __launch_bounds__(256) fro2_kernel(const JacobiProblem* __restrict__ probs, double* __restrict__ fro2) {
  const JacobiProblem P = probs[blockIdx.y];
  const size_t total = (size_t)P.M * P.N;   // ldg == M
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const double2 v = P.G[i];
    s += v.x * v.x + v.y * v.y;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && sh[0] != 0.0) atomicAdd(fro2 + blockIdx.y, sh[0]);
}

// After a sweep: a matrix that went through it without a rotation is done.  Also compacts the matrices still rotating into
// active[1..] (active[0] = their number) so that the next sweep's task space holds no slots of finished matrices: the last
// sweeps of a layer are run by one or two stragglers, and walking through the dead slots of the other matrices used to cost
// every CTA one global atomic per slot.
__global__ void __launch_bounds__(256) jacobi_check_kernel(int batch, int* dirty, int* done, int* remaining, int* active) {
  __shared__ int s_cnt[256];
  __shared__ int s_off[257];
  const int tid = threadIdx.x;
  const int per = (batch + 255) / 256;
  const int lo = min(batch, tid * per), hi = min(batch, lo + per);
  int c = 0;
  for (int m = lo; m < hi; ++m) {
    if (!done[m]) {
      if (!dirty[m]) done[m] = 1;
      else ++c;
    }
    dirty[m] = 0;
  }
  s_cnt[tid] = c;
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int i = 0; i < 256; ++i) { s_off[i] = o; o += s_cnt[i]; }
    s_off[256] = o;
    *remaining = o;
    if (active) active[0] = o;
  }
  __syncthreads();
  if (active) {
    int o = s_off[tid];
    for (int m = lo; m < hi; ++m)
      if (!done[m]) active[1 + o++] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// column norms -> descending order -> truncation rule -> write-back scales
__global__ void __launch_bounds__(256) trunc_kernel(const TruncProblem* __restrict__ probs, double cutoff, int cutoff_on_sqrt,
                                                    int max_bond, int gauge, int renorm, double null_tol) {
  const TruncProblem P = probs[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = P.M, N = P.N;
  for (int col = warp; col < N; col += 8) {
    const double2* g = P.G + (size_t)P.ldg * col;
    double s = 0.0;
    for (int r = lane; r < M; r += 32) { const double2 v = g[r]; s += v.x * v.x + v.y * v.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) P.sig2[col] = s;
  }
  __syncthreads();
  for (int k = tid; k < N; k += 256) {
    const double v = P.sig2[k];
    int rank = 0;
    for (int j = 0; j < N; ++j) {
      const double w = P.sig2[j];
      rank += (w > v || (w == v && j < k)) ? 1 : 0;
    }
    P.perm[rank] = k;
    P.sigma[rank] = sqrt(v);
  }
  __syncthreads();
  __shared__ int s_keep;
  __shared__ double s_rn;
  // numerically-null singular values (sigma_k <= null_tol * ||theta||_F, the threshold below which the sweeps leave a column
  // alone) stand for exact zeros: report them as 0.0 so that the reference's cut rule below sees what an exact SVD would
  // have returned
  __shared__ double s_fro;
  if (tid == 0) {
    double t = 0.0;
    for (int k = 0; k < N; ++k) t += P.sigma[k] * P.sigma[k];
    s_fro = sqrt(t);
  }
  __syncthreads();
  for (int k = tid; k < N; k += 256)
    if (k > 0 && P.sigma[k] <= null_tol * s_fro) P.sigma[k] = 0.0;
  __syncthreads();
  if (tid == 0) {
    // truncateSvdTensors, ExaTnMpsVisitor.cpp:2434-2445: first k whose partial norm is below eps, PLUS ONE
    int cut = N;
    for (int k = 0; k < N; ++k) {
      const double metric = cutoff_on_sqrt ? sqrt(P.sigma[k]) : P.sigma[k];
      if (metric < cutoff) { cut = k + 1; break; }
    }
    int keep = cut < max_bond ? cut : max_bond;
    if (keep < 1) keep = 1;
    double tot = 0.0, kept = 0.0;
    for (int k = 0; k < N; ++k) {
      const double s2 = P.sigma[k] * P.sigma[k];
      tot += s2;
      if (k < keep) kept += s2;
    }
    *P.keep = keep;
    P.weights[0] = tot;
    P.weights[1] = kept;
    s_keep = keep;
    s_rn = (renorm && kept > 0.0) ? sqrt(tot / kept) : 1.0;
  }
  __syncthreads();
  const int keep = s_keep;
  for (int k = tid; k < keep; k += 256) {
    const double s = P.sigma[k];
    double sP = 0.0, sO = 0.0;
    // sigma_k <= null_tol * sigma_max is numerically null: without an accumulated V its right vector is noise amplified by
    // tol * sigma_max / sigma_k, so both factors of that component are dropped (contribution to theta <= null_tol * sigma_max)
    if (s > 1e-100 && s > null_tol * s_fro) {
      // lo site carries sigma^el, hi site sigma^eh
      const double lo_f = (gauge == 0) ? sqrt(s) : (gauge == 1 ? 1.0 : s);
      const double hi_f = ((gauge == 0) ? sqrt(s) : (gauge == 1 ? s : 1.0)) * s_rn;
      if (P.tall) { sP = lo_f / s; sO = hi_f / (s * s); }
      else { sP = hi_f / s; sO = lo_f / (s * s); }
    }
    P.scaleP[k] = sP;
    P.scaleO[k] = sO;
  }
}

__global__ void gather_kernel(const GatherProblem* __restrict__ probs) {
  const GatherProblem P = probs[blockIdx.z];
  const int k = blockIdx.y;
  if (k >= P.keep) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.M) return;
  const double sc = P.scale[k];
  const double2 v = P.G[(size_t)i + (size_t)P.ldg * P.perm[k]];
  if (P.conjT) P.out[(size_t)k + (size_t)P.ldo * i] = make_double2(sc * v.x, -sc * v.y);
  else P.out[(size_t)i + (size_t)P.ldo * k] = make_double2(sc * v.x, sc * v.y);
}

}  // namespace

void launch_jacobi_sweep(const JacobiProblem* d_probs, int batch, int max_pairs, int nsteps, int base, double tol2, double dead2,
                         const double* d_fro2, int* d_dirty, const int* d_done, int* d_progress, int progress_stride, int* d_counter,
                         int* d_fault, int grid_ctas, int warps_per_task, const int* d_active, cudaStream_t s) {
  if (batch <= 0) return;
  const long total = (long)nsteps * batch * max_pairs;   // batch = upper bound of the matrices still rotating (d_active[0] on the device)
  const int grid = (int)std::min<long>(total, grid_ctas);
  if (warps_per_task == 16)
    jacobi_sweep_kernel<16><<<grid, 512, 0, s>>>(d_probs, batch, max_pairs, nsteps, base, tol2, dead2, d_fro2, d_dirty, d_done,
                                                d_progress, progress_stride, d_counter, d_fault, d_active);
  else if (warps_per_task == 8)
    jacobi_sweep_kernel<8><<<grid, 256, 0, s>>>(d_probs, batch, max_pairs, nsteps, base, tol2, dead2, d_fro2, d_dirty, d_done,
                                               d_progress, progress_stride, d_counter, d_fault, d_active);
  else
    jacobi_sweep_kernel<4><<<grid, JT, 0, s>>>(d_probs, batch, max_pairs, nsteps, base, tol2, dead2, d_fro2, d_dirty, d_done,
                                              d_progress, progress_stride, d_counter, d_fault, d_active);
}
double jacobi_dmma_flops() {
  unsigned long long v = 0;
  cudaMemcpyFromSymbol(&v, g_dmma_flops, sizeof(v));
  return (double)v;
}
void jacobi_set_debug_mode(int mode) { cudaMemcpyToSymbol(g_dbg_mode, &mode, sizeof(int)); }
void jacobi_print_phase_timing() {
  unsigned long long h[8];
  if (cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h)) != cudaSuccess || h[5] == 0) return;
  const double n = (double)h[5], nr = (double)(h[6] ? h[6] : 1);
  fprintf(stderr, "[mps_b200 phase timing] tasks %.0f (rotating %.0f): A %.0f cyc, reduce+test %.0f, B %.0f (per rotating), C %.0f (per rotating), dep-wait %.0f\n",
          n, (double)h[6], h[0] / n, h[1] / n, h[2] / nr, h[3] / nr, h[4] / n);
}
void launch_fro2(const JacobiProblem* d_probs, int batch, double* d_fro2, cudaStream_t s) {
  if (batch <= 0) return;
  dim3 grid(16, batch);
  fro2_kernel<<<grid, 256, 0, s>>>(d_probs, d_fro2);
}
void launch_jacobi_check(int batch, int* d_dirty, int* d_done, int* d_remaining, int* d_active, cudaStream_t s) {
  jacobi_check_kernel<<<1, 256, 0, s>>>(batch, d_dirty, d_done, d_remaining, d_active);
}
void launch_trunc(const TruncProblem* d_probs, int batch, double cutoff, int cutoff_on_sqrt, int max_bond, int gauge, int renorm,
                  double null_tol, cudaStream_t s) {
  if (batch <= 0) return;
  trunc_kernel<<<batch, 256, 0, s>>>(d_probs, cutoff, cutoff_on_sqrt, max_bond, gauge, renorm, null_tol);
}
void launch_gather(const GatherProblem* d_probs, int batch, int max_rows, int max_keep, cudaStream_t s) {
  if (batch <= 0 || max_rows <= 0 || max_keep <= 0) return;
  dim3 grid((max_rows + 255) / 256, max_keep, batch);
  gather_kernel<<<grid, 256, 0, s>>>(d_probs);
}

}  // namespace mpsb200
