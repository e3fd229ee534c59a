// Batched complex128 GEMM for the MPS engine: FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64) fed from
// shared-memory tiles staged by the TMA engine's bulk copies (cp.async.bulk -> SASS UBLKCP) with an
// mbarrier full/empty pipeline and a dedicated producer warp.  Blackwell's tcgen05 has no f64 kind, so
// DMMA + bulk-async staging is the sm_100a-native path for this arithmetic.
//
// Replaces, for the reference's hot path (ExaTnMpsVisitor.cpp):
//   :1412-1440  merge   D = Q_lo * Q_hi           (contractTensorsSync)
//   :1463-1524  gate    Result = D * G            (contractTensorsSync)         } "theta" mode: one kernel,
//   :1526-1541  host round trip of theta                                        } gate applied in registers
// and serves the V^H / U back-multiplication of the write-back and the transfer-matrix sweeps.
#include "kernels.h"
#include "ptx.cuh"

namespace mpsb200 {

namespace {
constexpr int TM = 64, TN = 64, TK = 16;
constexpr int LD_MC = 66;    // stride of an M-contiguous tile  S[k][m]  (== 2 mod 8 complex -> conflict-free frags)
constexpr int LD_KC = 20;    // stride of a  K-contiguous tile  S[m][k]  (== 4 mod 8 complex)
constexpr int TILE = 1280;   // complex elements per operand tile: max(16*66, 64*20)
constexpr int STAGES = 2;
constexpr int NTHREADS = 160;   // 4 DMMA warps (2x2, 32x32 complex each) + 1 producer warp
constexpr int SMEM_BYTES = STAGES * 2 * TILE * 16 + 16 * 16 + 64;

template <bool KC>
__device__ __forceinline__ double2 frag(const double2* S, int base, int k0, int lane) {
  return KC ? S[(base + (lane >> 2)) * LD_KC + k0 + (lane & 3)] : S[(k0 + (lane & 3)) * LD_MC + base + (lane >> 2)];
}

template <int LAYOUT>
__device__ __forceinline__ void zgemm_dmma_body(const GemmProblem* __restrict__ P) {
  constexpr bool AK = (LAYOUT == 1);
  constexpr bool BK = (LAYOUT != 2);
  const int M = P->M, N = P->N, K = P->K, mode = P->mode;
  const int tsz = mode ? 32 : 64;
  const int tiles_m = (M + tsz - 1) / tsz, tiles_n = (N + tsz - 1) / tsz;
  const int t = blockIdx.x;
  if (t >= tiles_m * tiles_n) return;   // uniform per CTA
  const int i0 = (t % tiles_m) * tsz, j0 = (t / tiles_m) * tsz;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* sA = reinterpret_cast<double2*>(smem_raw);
  double2* sB = sA + STAGES * TILE;
  double2* sGate = sB + STAGES * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sGate + 16);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 4);
    }
    fence_mbar_init();
  }
  if (mode && tid < 16) sGate[tid] = P->gate[tid];
  __syncthreads();

  const int KT = (K + TK - 1) / TK;
  const int lda = P->lda, ldb = P->ldb;

  if (warp == 4) {
    // ------------------------------------------------------------------ producer warp
    const double2* A = P->A;
    const double2* B = P->B;
    const int* ag = P->a_gather;
    const int* bg = P->b_gather;
    const int bcs = P->b_col_stride, bco = P->b_col_off;
    const int rows_valid = min(tsz, M - i0), cols_valid = min(tsz, N - j0);
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt & 1;
      if (kt >= STAGES) mbar_wait(empty0 + 8 * s, ((kt >> 1) - 1) & 1);
      const int k0 = kt * TK, klen = min(TK, K - k0);
      double2* As = sA + s * TILE;
      double2* Bs = sB + s * TILE;
      if (klen < TK) {   // zero the K tail of both tiles (generic proxy; published by the arrive below)
        const double2 z = make_double2(0.0, 0.0);
        for (int r = lane; r < 64; r += 32)
          for (int k = klen; k < TK; ++k) {
            if (AK) As[r * LD_KC + k] = z; else As[k * LD_MC + r] = z;
            if (BK) Bs[r * LD_KC + k] = z; else Bs[k * LD_MC + r] = z;
          }
      }
      __syncwarp();
      uint32_t bytes;
      {
        const uint32_t a_b = AK ? rows_valid * klen : klen * (mode ? 2 : 1) * rows_valid;
        const uint32_t b_b = BK ? (mode ? 2 : 1) * cols_valid * klen : klen * cols_valid;
        bytes = (a_b + b_b) * 16u;
      }
      const uint32_t fb = full0 + 8 * s;
      if (lane == 0) mbar_expect_tx(fb, bytes);
      __syncwarp();
      // ---- A tile
      if (AK) {
        for (int r = lane; r < rows_valid; r += 32) {
          const int g = ag ? ag[i0 + r] : (i0 + r);
          bulk_g2s(smem_u32(As + r * LD_KC), A + k0 + (size_t)lda * g, klen * 16, fb);
        }
      } else {
        const int nseg = mode ? 2 : 1;
        for (int u = lane; u < klen * nseg; u += 32) {
          const int k = u / nseg, sg = u - k * nseg;
          bulk_g2s(smem_u32(As + k * LD_MC + sg * 32), A + (size_t)sg * M + i0 + (size_t)lda * (k0 + k), rows_valid * 16, fb);
        }
      }
      // ---- B tile
      if (BK) {
        if (mode) {
          for (int c = lane; c < 64; c += 32) {
            const int q = c >> 5, cc = c & 31;
            if (cc < cols_valid)
              bulk_g2s(smem_u32(Bs + c * LD_KC), B + k0 + (size_t)ldb * (q + 2 * (j0 + cc)), klen * 16, fb);
          }
        } else {
          for (int c = lane; c < cols_valid; c += 32) {
            const int col = bg ? bg[j0 + c] : ((j0 + c) * bcs + bco);
            bulk_g2s(smem_u32(Bs + c * LD_KC), B + k0 + (size_t)ldb * col, klen * 16, fb);
          }
        }
      } else {
        for (int k = lane; k < klen; k += 32)
          bulk_g2s(smem_u32(Bs + k * LD_MC), B + j0 + (size_t)ldb * (k0 + k), cols_valid * 16, fb);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- DMMA warps
  const int wr = warp >> 1, wc = warp & 1;
  double accr[4][4][2], acci[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { accr[i][j][0] = accr[i][j][1] = acci[i][j][0] = acci[i][j][1] = 0.0; }

  constexpr double SA = AK ? -1.0 : 1.0;   // A enters conjugated in the "CN" layout
  constexpr double SB = BK ? 1.0 : -1.0;   // B enters conjugated in the "NC" layout
  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt & 1;
    mbar_wait(full0 + 8 * s, (kt >> 1) & 1);
    const double2* As = sA + s * TILE;
    const double2* Bs = sB + s * TILE;
    const int klen = min(TK, K - kt * TK);
    const int kmax = (klen + 3) & ~3;
#pragma unroll 1
    for (int k0 = 0; k0 < kmax; k0 += 4) {
      double2 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = frag<AK>(As, (i >> 1) * 32 + wr * 16 + (i & 1) * 8, k0, lane);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = frag<BK>(Bs, (j >> 1) * 32 + wc * 16 + (j & 1) * 8, k0, lane);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double ar = a[i].x;
        const double y1 = -(SA * SB) * a[i].y;   // multiplies b.im into Re
        const double y2 = SA * a[i].y;           // multiplies b.re into Im
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double br = b[j].x, bi = b[j].y, z = SB * b[j].y;
          dmma884(accr[i][j][0], accr[i][j][1], ar, br);
          dmma884(acci[i][j][0], acci[i][j][1], ar, z);
          dmma884(accr[i][j][0], accr[i][j][1], y1, bi);
          dmma884(acci[i][j][0], acci[i][j][1], y2, br);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
  }

  // -------------------------------------------------------------------- epilogue
  double2* C = P->C;
  double2* C2 = P->C2;
  const int ldc = P->ldc;
  const bool cT = P->conjT_out != 0;
  if (mode) {
    // theta(a,p,q,c) = sum_{p',q'} gate[2p+q][2p'+q'] * D_{p'q'}(a,c)
#pragma unroll
    for (int i2 = 0; i2 < 2; ++i2)
#pragma unroll
      for (int j2 = 0; j2 < 2; ++j2)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = i0 + wr * 16 + i2 * 8 + (lane >> 2);
          const int c = j0 + wc * 16 + j2 * 8 + 2 * (lane & 3) + e;
          if (a < M && c < N) {
            double vr[4], vi[4];
#pragma unroll
            for (int pq = 0; pq < 4; ++pq) {
              vr[pq] = accr[i2 + 2 * (pq >> 1)][j2 + 2 * (pq & 1)][e];
              vi[pq] = acci[i2 + 2 * (pq >> 1)][j2 + 2 * (pq & 1)][e];
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              double re = 0.0, im = 0.0;
#pragma unroll
              for (int pq = 0; pq < 4; ++pq) {
                const double2 g = sGate[o * 4 + pq];
                re += g.x * vr[pq] - g.y * vi[pq];
                im += g.x * vi[pq] + g.y * vr[pq];
              }
              const int row = a + M * (o >> 1), col = (o & 1) + 2 * c;
              const size_t idx = cT ? ((size_t)col + (size_t)ldc * row) : ((size_t)row + (size_t)ldc * col);
              // stabilizeTensorBody (ExaTnMpsVisitor.cpp:141-161): elements of the merged tensor with |x| < 1e-100 are
              // set to zero before the SVD
              if (re * re + im * im < 1e-200) { re = 0.0; im = 0.0; }
              const double2 v = make_double2(re, cT ? -im : im);
              C[idx] = v;
              if (C2) C2[idx] = v;
            }
          }
        }
  } else {
    const double alpha = P->alpha;
    const double* rs = P->row_scale;
    const double* cs = P->col_scale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i0 + (i >> 1) * 32 + wr * 16 + (i & 1) * 8 + (lane >> 2);
      if (row >= M) continue;
      const double fr = alpha * (rs ? rs[row] : 1.0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = j0 + (j >> 1) * 32 + wc * 16 + (j & 1) * 8 + 2 * (lane & 3) + e;
          if (col < N) {
            const double f = fr * (cs ? cs[col] : 1.0);
            const size_t idx = cT ? ((size_t)col + (size_t)ldc * row) : ((size_t)row + (size_t)ldc * col);
            const double2 v = make_double2(f * accr[i][j][e], cT ? -f * acci[i][j][e] : f * acci[i][j][e]);
            C[idx] = v;
            if (C2) C2[idx] = v;
          }
        }
    }
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(NTHREADS, 2) zgemm_dmma_kernel(const GemmProblem* __restrict__ probs) {
  zgemm_dmma_body<LAYOUT>(probs + blockIdx.y);
}
// single problem passed by value in the kernel parameter space (no descriptor upload)
template <int LAYOUT>
__global__ void __launch_bounds__(NTHREADS, 2) zgemm_dmma_kernel1(const __grid_constant__ GemmProblem prob) {
  zgemm_dmma_body<LAYOUT>(&prob);
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// Small-problem variant for the transfer-matrix sweeps (one chi x chi environment times one site): a 64x64-tiled launch
// of a 256 x 256 x 512 product is 16 CTAs on 148 SMs.  Here a CTA owns a 32x32 tile (4 warps, 16x16 complex each) and the
// DMMA fragments are loaded straight from global memory (L2-resident operands, 64-128 byte segments), K unrolled by four
// chunks so 16 loads are in flight per thread.  Plain products only: C = alpha * opA(A) * opB(B).
template <int LAYOUT>
__global__ void __launch_bounds__(128, 4) zgemm_small_kernel(const __grid_constant__ GemmProblem P) {
  constexpr bool AK = (LAYOUT == 1);   // A given as K x M, used conj-transposed
  constexpr bool BC = (LAYOUT == 2);   // B given as N x K, used conj-transposed
  const int M = P.M, N = P.N, K = P.K;
  const int tiles_m = (M + 31) / 32;
  const int i0 = (blockIdx.x % tiles_m) * 32, j0 = (blockIdx.x / tiles_m) * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lk = lane & 3;
  const int m0 = i0 + (warp >> 1) * 16, n0 = j0 + (warp & 1) * 16;
  const double2* __restrict__ A = P.A;
  const double2* __restrict__ B = P.B;
  const int lda = P.lda, ldb = P.ldb, bcs = P.b_col_stride, bco = P.b_col_off;
  double cre[2][2][2], cim[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }
  const double2 z = make_double2(0.0, 0.0);
  constexpr int UN = 4;
  for (int k0 = 0; k0 < K; k0 += 4 * UN) {
    double2 a[UN][2], b[UN][2];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int k = k0 + 4 * u + lk;
      const bool kok = k < K;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = m0 + 8 * i + lr;
        a[u][i] = (kok && m < M) ? (AK ? A[k + (size_t)lda * m] : A[m + (size_t)lda * k]) : z;
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = n0 + 8 * j + lr;
        b[u][j] = (kok && n < N) ? (BC ? B[n + (size_t)ldb * k] : B[k + (size_t)ldb * ((size_t)n * bcs + bco)]) : z;
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double ar = a[u][i].x, ai = AK ? -a[u][i].y : a[u][i].y;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double br = b[u][j].x, bi = BC ? -b[u][j].y : b[u][j].y;
          dmma884(cre[i][j][0], cre[i][j][1], ar, br);
          dmma884(cre[i][j][0], cre[i][j][1], -ai, bi);
          dmma884(cim[i][j][0], cim[i][j][1], ar, bi);
          dmma884(cim[i][j][0], cim[i][j][1], ai, br);
        }
      }
  }
  const double alpha = P.alpha;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = m0 + 8 * i + lr;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n0 + 8 * j + 2 * lk + e;
        if (col < N) P.C[row + (size_t)P.ldc * col] = make_double2(alpha * cre[i][j][e], alpha * cim[i][j][e]);
      }
  }
}

int gemm_tiles(int M, int N, int mode) {
  const int tsz = mode ? 32 : 64;
  return ((M + tsz - 1) / tsz) * ((N + tsz - 1) / tsz);
}

static void set_attrs() {
  // function attributes belong to a device context: once per device (site-sharded handles drive several from one process,
  // one host thread each; setting the same value twice from two threads is harmless)
  static bool attr_set_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr_set = attr_set_dev[dev & 63];
  if (attr_set) return;
  cudaFuncSetAttribute(zgemm_dmma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(zgemm_dmma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(zgemm_dmma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(zgemm_dmma_kernel1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(zgemm_dmma_kernel1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  cudaFuncSetAttribute(zgemm_dmma_kernel1<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  attr_set = true;
}

static bool g_small_gemm = true;
void gemm_set_small_path(int on) { g_small_gemm = on != 0; }

void launch_gemm1(const GemmProblem& p, int layout, cudaStream_t s) {
  const int tiles = gemm_tiles(p.M, p.N, p.mode);
  if (tiles <= 0) return;
  // too few 64x64 tiles to fill the GPU: 32x32 tiles, fragments straight from L2
  if (g_small_gemm && tiles < 120 && p.mode == 0 && !p.a_gather && !p.b_gather && !p.row_scale && !p.col_scale && !p.C2 && !p.conjT_out &&
      (layout == 0 || (p.b_col_stride == 1 && p.b_col_off == 0))) {
    const int t32 = ((p.M + 31) / 32) * ((p.N + 31) / 32);
    if (layout == 0) zgemm_small_kernel<0><<<t32, 128, 0, s>>>(p);
    else if (layout == 1) zgemm_small_kernel<1><<<t32, 128, 0, s>>>(p);
    else zgemm_small_kernel<2><<<t32, 128, 0, s>>>(p);
    return;
  }
  set_attrs();
  dim3 grid(tiles, 1);
  if (layout == 0) zgemm_dmma_kernel1<0><<<grid, NTHREADS, SMEM_BYTES, s>>>(p);
  else if (layout == 1) zgemm_dmma_kernel1<1><<<grid, NTHREADS, SMEM_BYTES, s>>>(p);
  else zgemm_dmma_kernel1<2><<<grid, NTHREADS, SMEM_BYTES, s>>>(p);
}

void launch_gemm(const GemmProblem* d_probs, int batch, int max_tiles, int layout, cudaStream_t s) {
  if (batch <= 0 || max_tiles <= 0) return;
  set_attrs();
  dim3 grid(max_tiles, batch);
  if (layout == 0) zgemm_dmma_kernel<0><<<grid, NTHREADS, SMEM_BYTES, s>>>(d_probs);
  else if (layout == 1) zgemm_dmma_kernel<1><<<grid, NTHREADS, SMEM_BYTES, s>>>(d_probs);
  else zgemm_dmma_kernel<2><<<grid, NTHREADS, SMEM_BYTES, s>>>(d_probs);
}

}  // namespace mpsb200
