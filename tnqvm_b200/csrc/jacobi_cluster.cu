// Block-Jacobi sweep with shared-memory-resident pair tasks split by rows over a thread-block cluster (sm_100a).
//
// Same algorithm, same tournament, same rotation rule and same global data (G, travelling Gram blocks, versions, clean-pair
// memo) as jacobi_sweep_kernel in jacobi_svd.cu -- what changes is where the columns live while a pair task runs:
//
//   * a pair task (two 8-column blocks of one matrix) is executed by a CLUSTER of CS CTAs; CTA `rank` owns the row slab
//     [rank * rpc, (rank + 1) * rpc) of the 16 columns and brings it into shared memory ONCE with 16 bulk copies on the TMA
//     engine (cp.async.bulk, one mbarrier), instead of streaming the columns from L2 twice (Gram, then update);
//   * phase A: each CTA forms the Gram partial of its slab on the FP64 tensor cores (DMMA) from shared memory; the leader
//     (rank 0) sums the partials of its peers through distributed shared memory after ONE cluster barrier;
//   * phase B: the leader alone runs the 16x16 rotation phase and leaves the scaled rotation parameters in its shared memory;
//     after the second cluster barrier the peers copy them (4 KB over DSMEM);
//   * phase C: every CTA rotates its slab (one thread per row, read from shared memory) and writes the rows straight back to
//     global memory with coalesced stores.
//
// L2 traffic per task: 2 x 16 columns (one read, one write) instead of 3, issued as 2-4 KB bulk transfers, and a lone matrix
// (layers of one to eight gates in routed circuits, the straggler tail of every layer, the per-device share of a sharded
// chain) keeps CS x more SMs busy.  Tasks are assigned to clusters round-robin in (step, matrix, pair) order; a task waits on
// per-(block, rank) progress flags exactly like the streaming kernel, so every cluster must be co-resident (the launcher
// sizes the grid with cudaOccupancyMaxActiveClusters).
#include "kernels.h"
#include "ptx.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>

namespace cg = cooperative_groups;

namespace mpsb200 {
namespace {

constexpr int CT = 128;      // threads per CTA
constexpr int WLD = 17;      // padded leading dimension of the 16x16 shared matrices
constexpr int MAXROT = 120;  // 8 cross rounds (+ 7 in-block rounds at the first step of a tournament) x 8 rotations
constexpr int MAXCS = 8;     // progress flags per block: one per cluster rank

__device__ unsigned long long g_flops_cluster = 0;
__device__ int g_cdbg = 0;                        // developer switch: 10 = per-phase clock64 accounting (thread 0 of every CTA)
__device__ unsigned long long g_cphase[2][12];    // [leader / peer][phase]: deps, load, gram, sync1, leader, sync2, params, apply, store, tasks, rotating tasks   // FP64 flops executed by the pair tasks of this kernel (reporting only)

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 rmul(double r, double2 a) { return make_double2(r * a.x, r * a.y); }

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

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// shared -> global bulk copy on the TMA engine (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct __align__(16) LeaderOut {   // what the leader hands to its peers after the rotation phase
  double2 tp[MAXROT], tq[MAXROT];
  double gam[16];
  int need;
  int pad[3];
};

// phase C: the rotations of one task applied to the rows of the slab in shared memory (scaled "fast Givens" form, see
// jacobi_svd.cu); one thread per row, the 16 entries of the row in registers.  Kept out of line so that its 64 data registers
// do not compete with the state the task loop keeps live.
__device__ __noinline__ void apply_rotations(const double2* slab, int ld, int nrows, int nA, int nB, const LeaderOut* out, bool within,
                                                double2* gA, double2* gB, int ldg) {
  const int tid = threadIdx.x;
  const LeaderOut& s_out = *out;
  for (int row = tid; row < nrows; row += CT) {
    // the rotation parameters are read from shared memory (broadcast) at their point of use: without this barrier the compiler
    // hoists all 240 loop-invariant loads out of the row loop and parks them in per-thread local memory
    asm volatile("" ::: "memory");
    double2 x[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = (j < nA) ? slab[(size_t)j * ld + row] : make_double2(0.0, 0.0);
      x[8 + j] = (j < nB) ? slab[(size_t)(8 + j) * ld + row] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int p = i, q = 8 + ((i + r) & 7);
        const double2 tp = s_out.tp[r * 8 + i], tq = s_out.tq[r * 8 + i];
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
          if (p > q) { const int tt = p; p = q; q = tt; }
          p += off; q += off;
          const double2 tp = s_out.tp[(8 + w) * 8 + i], tq = s_out.tq[(8 + w) * 8 + i];
          const double2 xp = x[p], xq = x[q];
          x[p] = make_double2(fma(-tp.x, xq.x, fma(tp.y, xq.y, xp.x)), fma(-tp.x, xq.y, fma(-tp.y, xq.x, xp.y)));
          x[q] = make_double2(fma(tq.x, xp.x, fma(-tq.y, xp.y, xq.x)), fma(tq.x, xp.y, fma(tq.y, xp.x, xq.y)));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double ga = s_out.gam[j], gb = s_out.gam[8 + j];
      // straight to global memory (coalesced over the rows of a warp): plain stores retire asynchronously, a bulk store from
      // shared memory would have to be waited for before the progress flags may be published
      if (j < nA) gA[(size_t)ldg * j + row] = make_double2(ga * x[j].x, ga * x[j].y);
      if (j < nB) gB[(size_t)ldg * j + row] = make_double2(gb * x[8 + j].x, gb * x[8 + j].y);
    }
  }
}

// One sweep (nsteps tournament steps) of the matrices listed in `active`.  Launch: clusters of CS CTAs of CT threads, dynamic
// shared memory = 16 * (rpc_cap + 4) complex (the row slab, leading dimension = 4 mod 8 so that the DMMA fragment loads of a
// quarter-warp fall on distinct banks).
__global__ void __launch_bounds__(CT, 4) jacobi_cluster_sweep_kernel(const JacobiProblem* __restrict__ probs, int batch, int max_pairs, int nsteps, int base,
                                                                 double tol2, double dead2, const double* __restrict__ fro2, int* __restrict__ dirty,
                                                                 const int* __restrict__ done, int* __restrict__ progress, int progress_stride,
                                                                 int* __restrict__ fault, const int* __restrict__ active, int rpc_cap) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ld = rpc_cap + 4;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* slab = reinterpret_cast<double2*>(smem_raw);   // [16][ld]
  __shared__ __align__(16) double s_acc[7 * 64];          // Gram partial of this CTA's slab (leader: of the whole task after the reduce)
  __shared__ double2 sW[16 * WLD];
  __shared__ LeaderOut s_out;                             // leader: written between the two cluster barriers; peers: their copy
  __shared__ double s_rc[8];
  __shared__ double2 s_rs[8];
  __shared__ int s_rp[8], s_rq[8];
  __shared__ double s_gam[16], s_igam[16];
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ int s_flag;

  const uint32_t mbar = smem_u32(&s_mbar);
  if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
  __syncthreads();
  uint32_t mphase = 0;

  if (active) batch = __ldcg(active);
  const int per_step = batch * max_pairs;
  const long total = (long)nsteps * per_step;

  for (long t = cluster_id; t < total; t += n_clusters) {
    const int step = (int)(t / per_step), r_ = (int)(t - (long)step * per_step);
    const int slot = r_ / max_pairs, pi = r_ - slot * max_pairs;
    const int mat = active ? __ldcg(active + 1 + slot) : slot;
    if (done[mat]) continue;
    const JacobiProblem P = probs[mat];
    const int npairs = (P.nb == 1) ? 1 : P.nbe / 2;
    if (pi >= npairs) continue;
    int blkA, blkB;
    bool within;
    if (!pair_blocks(P, step, pi, blkA, blkB, within)) continue;
    // every decision up to the first cluster barrier is taken from the same global words by all CTAs of the cluster

    const int M = P.M, N = P.N, ldg = P.ldg;
    // row slab of this rank: multiples of 4 rows (DMMA row quads)
    const int rpc = min(rpc_cap, ((((M + CS - 1) / CS) + 3) >> 2) << 2);
    const int r0 = min(M, rank * rpc), nrows = min(M, r0 + rpc) - r0;
    int* prog = progress + (size_t)mat * progress_stride;
    int* flagA = prog + blkA * MAXCS + rank;
    int* flagB = (blkB >= 0) ? prog + blkB * MAXCS + rank : nullptr;

    const bool timing = (g_cdbg == 10) && tid == 0;
    long long tk[10];
    if (timing) tk[0] = clock64();
    // ---- dependencies: this rank's slab of both blocks must have been published by the previous step's tasks
    if (tid == 0) {
      const int need = base + step;
      unsigned spins = 0;
      while (ld_acquire(flagA) < need && ++spins < (1u << 23)) __nanosleep(40);
      if (flagB)
        while (ld_acquire(flagB) < need && ++spins < (1u << 23)) __nanosleep(40);
      if (spins >= (1u << 23)) atomicAdd(fault, 1);
    }
    __syncthreads();
    const bool lone = (!within && blkB < 0);   // a lone block has no cross pairs
    // clean-pair memo (see jacobi_svd.cu): neither block has rotated since their cross pairs were last found orthogonal
    int verA = 0, verB = 0;
    int2* recp = nullptr;
    bool clean = false;
    if (!within && !lone) {
      verA = __ldcg(P.ver + blkA); verB = __ldcg(P.ver + blkB);
      recp = P.rec + (size_t)min(blkA, blkB) * P.nbe + max(blkA, blkB);
      const int2 rec = __ldcg(recp);
      const int va = blkA < blkB ? verA : verB, vb = blkA < blkB ? verB : verA;
      clean = (rec.x == va + 1 && rec.y == vb + 1);
    }
    if (lone || clean) {
      if (tid == 0) { st_release(flagA, base + step + 1); if (flagB) st_release(flagB, base + step + 1); }
      continue;
    }

    const int nA = min(8, N - blkA * 8), nB = (blkB >= 0) ? min(8, N - blkB * 8) : 0;
    if (timing) tk[1] = clock64();
    // ---- bring the slab in: one bulk copy per existing column
    if (tid == 0 && nrows > 0) {
      fence_async_all();   // the flags were acquired through the generic proxy; the copies below read global through the async proxy
      const uint32_t seg = (uint32_t)nrows * 16u;
      mbar_expect_tx(mbar, seg * (uint32_t)(nA + nB));
      const double2* gA = P.G + (size_t)ldg * (blkA * 8) + r0;
      for (int j = 0; j < nA; ++j) bulk_g2s(smem_u32(slab + (size_t)j * ld), gA + (size_t)ldg * j, seg, mbar);
      if (nB > 0) {
        const double2* gB = P.G + (size_t)ldg * (blkB * 8) + r0;
        for (int j = 0; j < nB; ++j) bulk_g2s(smem_u32(slab + (size_t)(8 + j) * ld), gB + (size_t)ldg * j, seg, mbar);
      }
    }
    if (tid < 16) { s_gam[tid] = 1.0; s_igam[tid] = 1.0; }
    for (int i = tid; i < 7 * 64; i += CT) s_acc[i] = 0.0;
    if (nrows > 0) { mbar_wait(mbar, mphase); mphase ^= 1; }
    __syncthreads();
    if (timing) tk[2] = clock64();

    // ---- phase A: Gram partial of the slab on DMMA, operands from shared memory
    {
      const int cslot = lane >> 2, rsub = lane & 3;
      const bool hasA = cslot < nA, hasB = cslot < nB;
      const double2* p0 = slab + (size_t)cslot * ld;
      const double2* p1 = slab + (size_t)(8 + cslot) * ld;
      double w00[2] = {0, 0}, m00[2] = {0, 0}, w11[2] = {0, 0}, m11[2] = {0, 0}, w01[2] = {0, 0}, p01[2] = {0, 0}, q01[2] = {0, 0};
      const int nch = (nrows + 3) >> 2;
      for (int ch = warp; ch < nch; ch += CT / 32) {
        const int row = 4 * ch + rsub;
        const bool ok = row < nrows;
        const double2 x0 = (ok && hasA) ? p0[row] : make_double2(0.0, 0.0);
        const double2 x1 = (ok && hasB) ? p1[row] : make_double2(0.0, 0.0);
        if (within) {
          dmma884(w00[0], w00[1], x0.x, x0.x);
          dmma884(m00[0], m00[1], x0.x, x0.y);
          dmma884(w11[0], w11[1], x1.x, x1.x);
          dmma884(m11[0], m11[1], x1.x, x1.y);
          dmma884(w00[0], w00[1], x0.y, x0.y);
          dmma884(w11[0], w11[1], x1.y, x1.y);
        }
        dmma884(w01[0], w01[1], x0.x, x1.x);
        dmma884(p01[0], p01[1], x0.x, x1.y);
        dmma884(q01[0], q01[1], x0.y, x1.x);
        dmma884(w01[0], w01[1], x0.y, x1.y);
      }
      // C fragment: element (row = lane>>2, col = 2*(lane&3)+e); the four warps add their fragments in a fixed order
      const int e0 = (lane >> 2) * 8 + 2 * (lane & 3);
      for (int w = 0; w < CT / 32; ++w) {
        if (warp == w) {
          if (within) {
            s_acc[0 * 64 + e0] += w00[0]; s_acc[0 * 64 + e0 + 1] += w00[1];
            s_acc[1 * 64 + e0] += m00[0]; s_acc[1 * 64 + e0 + 1] += m00[1];
            s_acc[2 * 64 + e0] += w11[0]; s_acc[2 * 64 + e0 + 1] += w11[1];
            s_acc[3 * 64 + e0] += m11[0]; s_acc[3 * 64 + e0 + 1] += m11[1];
          }
          s_acc[4 * 64 + e0] += w01[0]; s_acc[4 * 64 + e0 + 1] += w01[1];
          s_acc[5 * 64 + e0] += p01[0]; s_acc[5 * 64 + e0 + 1] += p01[1];
          s_acc[6 * 64 + e0] += q01[0]; s_acc[6 * 64 + e0 + 1] += q01[1];
        }
        __syncthreads();
      }
    }
    if (timing) tk[3] = clock64();
    cluster.sync();   // #1: every partial is in place
    if (timing) tk[4] = clock64();

    if (rank == 0) {
      // ---- leader: reduce over the cluster (fixed order), Gram matrix, convergence test, rotation phase
      for (int r = 1; r < CS; ++r) {
        const double* peer = cluster.map_shared_rank(s_acc, r);
        for (int i = tid + (within ? 0 : 4 * 64); i < 7 * 64; i += CT) s_acc[i] += peer[i];
      }
      if (tid == 0) s_flag = 0;
      __syncthreads();
      const double dead_abs = dead2 * fro2[mat];   // (null_tol * ||G||_F)^2: above the rounding-noise floor eps * sigma_max * sqrt(rotations) of a null column
      const double2* wdA = P.wd + (size_t)blkA * 64;
      const double2* wdB = P.wd + (size_t)(blkB >= 0 ? blkB : blkA) * 64;
      for (int i = tid; i < 256; i += CT) {
        const int p = i >> 4, q = i & 15;
        const int bp = p >> 3, bq = q >> 3, r = p & 7, c = q & 7;
        double re, im;
        if (bp == bq && !within) { const double2 v = __ldcg((bp ? wdB : wdA) + r * 8 + c); re = v.x; im = v.y; }
        else if (bp == 0 && bq == 0) { re = s_acc[0 * 64 + r * 8 + c]; im = s_acc[1 * 64 + r * 8 + c] - s_acc[1 * 64 + c * 8 + r]; }
        else if (bp == 1 && bq == 1) { re = s_acc[2 * 64 + r * 8 + c]; im = s_acc[3 * 64 + r * 8 + c] - s_acc[3 * 64 + c * 8 + r]; }
        else if (bp == 0) { re = s_acc[4 * 64 + r * 8 + c]; im = s_acc[5 * 64 + r * 8 + c] - s_acc[6 * 64 + r * 8 + c]; }
        else { re = s_acc[4 * 64 + c * 8 + r]; im = -(s_acc[5 * 64 + c * 8 + r] - s_acc[6 * 64 + c * 8 + r]); }
        sW[p * WLD + q] = make_double2(re, im);
      }
      __syncthreads();
      {
        int need = 0;
        for (int i = tid; i < 256; i += CT) {
          const int p = i >> 4, q = i & 15;
          if (p < q && (within || (p < 8 && q >= 8))) {
            const double a = sW[p * WLD + p].x, b = sW[q * WLD + q].x;
            const double2 g = sW[p * WLD + q];
            if (a > dead_abs && b > dead_abs && (g.x * g.x + g.y * g.y) > tol2 * a * b) need = 1;
          }
        }
        if (need) s_flag = 1;   // benign race: all writers store 1
      }
      __syncthreads();
      const int need = s_flag;
      const int nrounds = within ? 15 : 8;
      if (!need) {
        if (!within && tid == 0) {
          const int va = blkA < blkB ? verA : verB, vb = blkA < blkB ? verB : verA;
          *recp = make_int2(va + 1, vb + 1);
        }
        if (tid == 0) atomicAdd(&g_flops_cluster, (unsigned long long)((M + 3) >> 2) * (within ? 5120ull : 2048ull));
      } else {
        if (tid == 0) {
          dirty[mat] = 1;
          atomicAdd(&g_flops_cluster, (unsigned long long)((M + 3) >> 2) * (within ? 5120ull : 2048ull) + (unsigned long long)M * (unsigned long long)(2 * (nrounds * 64 + 32)));
        }
        // ---- phase B (same schedule and rotation formulas as jacobi_svd.cu): warps 0-1 apply, lanes 0-7 derive
        for (int r = 0; r < nrounds; ++r) {
          int rot = 0;
          if (tid < 8) {
            int p, q;
            if (r < 8) { p = tid; q = 8 + ((tid + r) & 7); }
            else {
              const int w = r - 8, l = tid & 3, off = (tid >> 2) * 8;
              p = (l == 0) ? 7 : (w + l) % 7;
              q = (w + 7 - l) % 7;
              if (p > q) { const int tt = p; p = q; q = tt; }
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
              const double d = b - a;
              const double rh = rsqrt(d * d + 4.0 * g2);
              const double x = 0.5 * (1.0 + fabs(d) * rh);
              const double rx = rsqrt(x);   // 1 / c
              c = x * rx;
              sg = rmul(d >= 0.0 ? rh * rx : -(rh * rx), g);
              const double2 tt = rmul(rx, sg);
              const double gp = s_gam[p], gq = s_gam[q], igp = s_igam[p], igq = s_igam[q];
              tp = rmul(gq * igp, make_double2(tt.x, -tt.y));
              tq = rmul(gp * igq, tt);
              s_gam[p] = gp * c; s_gam[q] = gq * c;
              s_igam[p] = igp * rx; s_igam[q] = igq * rx;
            }
            s_rc[tid] = c; s_rs[tid] = sg; s_rp[tid] = p; s_rq[tid] = q;
            s_out.tp[r * 8 + tid] = tp; s_out.tq[r * 8 + tid] = tq;
          }
          if (!__syncthreads_or(rot)) continue;
          if (warp < 2) {
            const int ia = tid >> 3, ib = tid & 7;
            const int pa = s_rp[ia], qa = s_rq[ia], pb = s_rp[ib], qb = s_rq[ib];
            const double ca = s_rc[ia], cb = s_rc[ib];
            const double2 sa = s_rs[ia], sb = s_rs[ib];
            const double2 w00 = sW[pa * WLD + pb], w01 = sW[pa * WLD + qb], w10 = sW[qa * WLD + pb], w11 = sW[qa * WLD + qb];
            const double2 t00 = csub(rmul(ca, w00), cmul(sa, w10));
            const double2 t01 = csub(rmul(ca, w01), cmul(sa, w11));
            const double2 t10 = cadd(cmulc(sa, w00), rmul(ca, w10));
            const double2 t11 = cadd(cmulc(sa, w01), rmul(ca, w11));
            sW[pa * WLD + pb] = csub(rmul(cb, t00), cmulc(sb, t01));
            sW[pa * WLD + qb] = cadd(cmul(sb, t00), rmul(cb, t01));
            sW[qa * WLD + pb] = csub(rmul(cb, t10), cmulc(sb, t11));
            sW[qa * WLD + qb] = cadd(cmul(sb, t10), rmul(cb, t11));
          }
          __syncthreads();
        }
        if (tid == 0) {   // this task is the only owner of both blocks during this step
          if (within) { verA = __ldcg(P.ver + blkA); if (blkB >= 0) verB = __ldcg(P.ver + blkB); }
          P.ver[blkA] = verA + 1;
          if (blkB >= 0) P.ver[blkB] = verB + 1;
        }
        if (tid < 16) s_out.gam[tid] = s_gam[tid];
      }
      if ((need || within) && tid < 128) {   // the travelling Gram blocks follow the rotations / are refreshed at a tournament's first step
        const int bb = tid >> 6, r = (tid >> 3) & 7, c = tid & 7;
        if (bb == 0 || blkB >= 0) P.wd[(size_t)(bb ? blkB : blkA) * 64 + r * 8 + c] = sW[(8 * bb + r) * WLD + 8 * bb + c];
      }
      if (tid == 0) s_out.need = need;
      __threadfence();   // ver / rec / wd / dirty before the peers publish their flags
    }
    if (timing) tk[5] = clock64();
    cluster.sync();   // #2: the leader's verdict and rotation parameters are in its shared memory
    if (timing) tk[6] = clock64();

    if (rank != 0) {
      const LeaderOut* lo = cluster.map_shared_rank(&s_out, 0);
      const int need = lo->need;
      if (need) {
        const int nrot = within ? MAXROT : 64;
        for (int i = tid; i < nrot; i += CT) { s_out.tp[i] = lo->tp[i]; s_out.tq[i] = lo->tq[i]; }
        if (tid < 16) s_out.gam[tid] = lo->gam[tid];
      }
      if (tid == 0) s_out.need = need;
    }
    __syncthreads();

    if (timing) tk[7] = clock64();
    if (s_out.need && nrows > 0) {
      apply_rotations(slab, ld, nrows, nA, nB, &s_out, within, P.G + (size_t)ldg * (blkA * 8) + r0,
                      P.G + (size_t)ldg * ((blkB >= 0 ? blkB : blkA) * 8) + r0, ldg);
      if (timing) tk[8] = clock64();
    }
    __threadfence();
    __syncthreads();   // every thread's stores are issued and fenced; thread 0 publishes them
    if (tid == 0) {
      st_release(flagA, base + step + 1);
      if (flagB) st_release(flagB, base + step + 1);
    }
    if (timing) {
      const long long te = clock64();
      unsigned long long* acc = g_cphase[rank ? 1 : 0];
      const bool rot = s_out.need && nrows > 0;
      atomicAdd(acc + 0, (unsigned long long)(tk[1] - tk[0]));
      atomicAdd(acc + 1, (unsigned long long)(tk[2] - tk[1]));
      atomicAdd(acc + 2, (unsigned long long)(tk[3] - tk[2]));
      atomicAdd(acc + 3, (unsigned long long)(tk[4] - tk[3]));
      atomicAdd(acc + 4, (unsigned long long)(tk[5] - tk[4]));
      atomicAdd(acc + 5, (unsigned long long)(tk[6] - tk[5]));
      atomicAdd(acc + 6, (unsigned long long)(tk[7] - tk[6]));
      if (rot) { atomicAdd(acc + 7, (unsigned long long)(tk[8] - tk[7])); atomicAdd(acc + 8, (unsigned long long)(te - tk[8])); atomicAdd(acc + 10, 1ull); }
      atomicAdd(acc + 9, 1ull);
    }
    __syncthreads();   // the slab and s_out may be overwritten by the next task from here on
  }
  cluster.sync();   // no CTA of a cluster exits while a peer may still read its shared memory
}

}  // namespace

int jacobi_cluster_progress_ints_per_block() { return MAXCS; }

// rows per CTA and cluster size for a chunk whose tallest matrix has max_m rows; returns false when the slab does not fit
bool jacobi_cluster_shape(int max_m, int* cs, int* rpc_cap) {
  int c = 1;
  while (c < MAXCS && (max_m + c - 1) / c > 128) c *= 2;
  int rpc = (((max_m + c - 1) / c + 7) >> 3) << 3;   // multiple of 8: leading dimension rpc + 4 = 4 mod 8
  if (rpc < 8) rpc = 8;
  const size_t smem = (size_t)16 * (rpc + 4) * sizeof(double2);
  if (smem > 200 * 1024) return false;
  *cs = c; *rpc_cap = rpc;
  return true;
}

// the dynamic shared-memory limit of the kernel only ever grows (per device context)
static bool ensure_smem_attr(size_t smem) {
  static size_t attr_smem[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > attr_smem[dev & 63]) {
    if (cudaFuncSetAttribute(jacobi_cluster_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
    attr_smem[dev & 63] = smem;
  }
  return true;
}

// returns the number of CTAs launched (0: nothing to do), -1 if the configuration cannot be launched
int launch_jacobi_cluster_sweep(const JacobiProblem* d_probs, int batch, int max_pairs, int nsteps, int base, double tol2, double dead2,
                                const double* d_fro2, int* d_dirty, const int* d_done, int* d_progress, int progress_stride, int* d_fault,
                                const int* d_active, int cs, int rpc_cap, int max_clusters, cudaStream_t s) {
  if (batch <= 0) return 0;
  const size_t smem = (size_t)16 * (rpc_cap + 4) * sizeof(double2);
  if (!ensure_smem_attr(smem)) return -1;
  const long total = (long)nsteps * batch * max_pairs;
  long nc = std::min<long>(total, max_clusters);
  if (nc < 1) nc = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nc * cs), 1, 1);
  cfg.blockDim = dim3(CT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, jacobi_cluster_sweep_kernel, d_probs, batch, max_pairs, nsteps, base, tol2, dead2, d_fro2, d_dirty, d_done, d_progress,
                         progress_stride, d_fault, d_active, rpc_cap) != cudaSuccess)
    return -1;
  return (int)(nc * cs);
}

// co-resident clusters of this kernel on the current device for the given shape (the dataflow waits need every cluster resident)
int jacobi_cluster_max_clusters(int cs, int rpc_cap) {
  const size_t smem = (size_t)16 * (rpc_cap + 4) * sizeof(double2);
  if (!ensure_smem_attr(smem)) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cs * 1024), 1, 1);
  cfg.blockDim = dim3(CT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, jacobi_cluster_sweep_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void jacobi_cluster_set_debug(int mode) { cudaMemcpyToSymbol(g_cdbg, &mode, sizeof(int)); }
void jacobi_cluster_print_phase_timing() {
  unsigned long long h[2][12];
  if (cudaMemcpyFromSymbol(h, g_cphase, sizeof(h)) != cudaSuccess) return;
  static const char* nm[9] = {"deps", "load", "gram", "sync1", "leader", "sync2", "params", "apply", "store"};
  for (int k = 0; k < 2; ++k) {
    if (h[k][9] == 0) continue;
    fprintf(stderr, "[mps_b200 cluster timing] %s CTAs: %llu tasks (%llu rotating), cycles per task:", k ? "peer" : "leader", h[k][9], h[k][10]);
    for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.0f", nm[i], (double)h[k][i] / (double)((i >= 7) ? (h[k][10] ? h[k][10] : 1) : h[k][9]));
    fprintf(stderr, "\n");
  }
}
double jacobi_cluster_dmma_flops() {
  unsigned long long v = 0;
  cudaMemcpyFromSymbol(&v, g_flops_cluster, sizeof(v));
  return (double)v;
}

}  // namespace mpsb200
