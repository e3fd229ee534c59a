// Batched blocked Householder QR (R factor only), complex128, sm_100a: the tensor-core pre-reduction of the
// two-qubit-gate SVD (the reference calls LAPACK zgesvd/zgesdd through exatn::decomposeTensorSVDLRSync,
// ExaTnMpsVisitor.cpp:1619-1627).
//
// theta_o (M x N, M >= N) = Q R.  Only R is kept: the one-sided Jacobi then runs on G = R^H (N x N), which
//   (a) shrinks the Jacobi rows from M to N when theta is rectangular, and
//   (b) preconditions it: (R^H)^H (R^H) = R R^H is one LR-Cholesky step closer to diagonal than theta^H theta, so
//       the number of Jacobi sweeps drops ~3x on the graded spectra of truncated MPS bonds.
// Q is never formed or applied: the singular vectors come back from theta itself (engine.cu write-back).
//
// Per panel of PB = 16 columns, two launches over the whole batch:
//   qr_panel_kernel   one CTA per matrix; the panel lives in shared memory, warp c owns column c; per column one
//                     reflector (zlarfg convention, H = I - tau v v^H) and its application to the later panel columns
//                     with warp-shuffle reductions; then the compact-WY factor T (zlarft, forward/columnwise).
//   qr_update_kernel  one CTA per 32 trailing columns: C <- (I - V T V^H)^H C = C - V (T^H (V^H C)), both products
//                     on the FP64 tensor cores (DMMA m8n8k4), complex arithmetic as 4 real MMAs.
#include "kernels.h"
#include "ptx.cuh"

namespace mpsb200 {
namespace {

constexpr int PB = QR_PB;            // panel width
constexpr int PT = 32 * PB;          // panel-kernel threads: one warp per panel column
constexpr int PANEL_SMEM_CAP = 208 * 1024;
constexpr int SLAB = 32;             // trailing columns per update CTA
constexpr int UT = 128;              // update-kernel threads

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(PT, 1) qr_panel_kernel(const QrProblem* __restrict__ probs, int k) {
  const QrProblem P = probs[blockIdx.x];
  const int c0 = k * PB;
  if (c0 >= P.N) return;
  const int pw = min(PB, P.N - c0);
  const int r0 = c0, mk = P.M - r0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sp = reinterpret_cast<double2*>(smem_raw);
  __shared__ double2 s_tau[PB];
  __shared__ double2 sS[PB][PB + 1];
  __shared__ double2 sT[PB][PB + 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double2* gp = P.Y + r0 + (size_t)P.ldy * c0;   // panel in global memory
  const bool in_smem = (size_t)mk * pw * sizeof(double2) <= (size_t)PANEL_SMEM_CAP;
  double2* Pn;
  int ld;
  if (in_smem) {
    Pn = sp; ld = mk;
    // the panel is pw columns of mk contiguous elements: every thread keeps four independent 16-byte loads in flight
    const int total = mk * pw;
    for (int e0 = tid * 4; e0 < total; e0 += PT * 4) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u;
        if (e < total) { const int c = e / mk, i = e - c * mk; v[u] = gp[i + (size_t)P.ldy * c]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u < total) sp[e0 + u] = v[u];
    }
  } else {
    Pn = gp; ld = P.ldy;
  }
  __syncthreads();

  // squared norm of the part of column 0 below the diagonal (each later column gets its own while it is updated)
  // The reflector vectors stay UNSCALED in the panel while it is factored (v_j = x_j * inv_j below the diagonal, 1 on it):
  // the warps that apply reflector j read column j while warp j only touches its diagonal, so there is no race and no
  // extra barrier; the scale is applied where V is written out.
  __shared__ double s_nrm[PB];
  __shared__ double2 s_inv[PB];
  if (warp == 0) {
    const double2* x = Pn;
    double s0 = 0.0, s1 = 0.0;
    int i = 1 + lane;
    for (; i + 32 < mk; i += 64) {
      const double2 a = x[i], b = x[i + 32];
      s0 += a.x * a.x + a.y * a.y; s1 += b.x * b.x + b.y * b.y;
    }
    if (i < mk) { const double2 a = x[i]; s0 += a.x * a.x + a.y * a.y; }
    s0 = warp_sum(s0 + s1);
    if (lane == 0) s_nrm[0] = s0;
  }
  __syncthreads();

  for (int j = 0; j < pw; ++j) {
    // every warp derives reflector j's scalars redundantly from column j's diagonal entry and tail norm (no extra barrier)
    const double2 alpha = Pn[j + (size_t)ld * j];
    const double s = s_nrm[j];
    double2 tau = make_double2(0.0, 0.0), inv = make_double2(0.0, 0.0);
    double beta = 0.0;
    const bool live = s > 0.0;
    if (live) {
      // zlarfg: beta = -sign(re alpha) * ||(alpha, x)||, tau = (beta - alpha)/beta, v = x / (alpha - beta)
      const double an = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + s);
      beta = alpha.x >= 0.0 ? -an : an;
      tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
      const double dx = alpha.x - beta, dy = alpha.y;
      const double dn = 1.0 / (dx * dx + dy * dy);
      inv = make_double2(dx * dn, -dy * dn);
    }
    if (tid == 0) { s_tau[j] = tau; s_inv[j] = inv; }
    __syncthreads();   // everybody has read alpha before warp j overwrites it
    if (warp == j) {
      if (live && lane == 0) Pn[j + (size_t)ld * j] = make_double2(beta, 0.0);
    } else if (warp > j && warp < pw && live) {
      // a <- H^H a = a - conj(tau) v (v^H a) with v = x * inv below the diagonal, v(j) = 1
      const double2* x = Pn + (size_t)ld * j;
      double2* a = Pn + (size_t)ld * warp;
      double wr0 = 0.0, wi0 = 0.0, wr1 = 0.0, wi1 = 0.0;
      int i = j + 1 + lane;
      for (; i + 32 < mk; i += 64) {
        const double2 t0 = cmulc(x[i], a[i]), t1 = cmulc(x[i + 32], a[i + 32]);
        wr0 += t0.x; wi0 += t0.y; wr1 += t1.x; wi1 += t1.y;
      }
      if (i < mk) { const double2 t0 = cmulc(x[i], a[i]); wr0 += t0.x; wi0 += t0.y; }
      const double xr = warp_sum(wr0 + wr1), xi = warp_sum(wi0 + wi1);
      // v^H a = conj(inv) * (x^H a) + a_j
      const double2 aj = a[j];
      const double2 xa = cmul(make_double2(inv.x, -inv.y), make_double2(xr, xi));
      const double2 w = make_double2(xa.x + aj.x, xa.y + aj.y);
      const double2 f = cmul(make_double2(tau.x, -tau.y), w);         // conj(tau) * (v^H a)
      const double2 fi = cmul(f, inv);                                 // a_i -= f * v_i = (f * inv) * x_i
      double n0 = 0.0, n1 = 0.0;
      const bool next = (warp == j + 1);                               // this column is the next reflector: take its tail norm now
      i = j + 1 + lane;
      for (; i + 32 < mk; i += 64) {
        const double2 u0 = cmul(fi, x[i]), u1 = cmul(fi, x[i + 32]);
        double2 b0 = a[i], b1 = a[i + 32];
        b0.x -= u0.x; b0.y -= u0.y; b1.x -= u1.x; b1.y -= u1.y;
        a[i] = b0; a[i + 32] = b1;
        if (next) { if (i > j + 1) n0 += b0.x * b0.x + b0.y * b0.y; n1 += b1.x * b1.x + b1.y * b1.y; }
      }
      if (i < mk) {
        const double2 u0 = cmul(fi, x[i]);
        double2 b0 = a[i];
        b0.x -= u0.x; b0.y -= u0.y;
        a[i] = b0;
        if (next && i > j + 1) n0 += b0.x * b0.x + b0.y * b0.y;
      }
      __syncwarp();   // every lane has read a[j] (aj) before lane 0 replaces it
      if (lane == 0) a[j] = make_double2(aj.x - f.x, aj.y - f.y);
      if (next) { n0 = warp_sum(n0 + n1); if (lane == 0) s_nrm[j + 1] = n0; }
    } else if (warp == j + 1 && warp < pw && !live) {
      // H_j = I: column j+1 is unchanged, its tail norm is still needed
      const double2* a = Pn + (size_t)ld * warp;
      double n0 = 0.0;
      for (int i = j + 2 + lane; i < mk; i += 32) { const double2 b0 = a[i]; n0 += b0.x * b0.x + b0.y * b0.y; }
      n0 = warp_sum(n0);
      if (lane == 0) s_nrm[j + 1] = n0;
    }
    __syncthreads();
  }

  // S = V^H V (strict upper part), V unit lower trapezoidal: warp b forms S(0:b, b) in one pass over its column
  if (warp < pw && warp > 0) {
    const int b = warp;
    const double2* vb = Pn + (size_t)ld * b;
    double sr[PB - 1], si[PB - 1];
#pragma unroll
    for (int a = 0; a < PB - 1; ++a) { sr[a] = 0.0; si[a] = 0.0; }
    for (int i = b + 1 + lane; i < mk; i += 32) {
      const double2 y = vb[i];
#pragma unroll
      for (int a = 0; a < PB - 1; ++a)
        if (a < b) { const double2 t = cmulc(Pn[i + (size_t)ld * a], y); sr[a] += t.x; si[a] += t.y; }
    }
#pragma unroll
    for (int a = 0; a < PB - 1; ++a)
      if (a < b) {
        const double r = warp_sum(sr[a]), im = warp_sum(si[a]);
        if (lane == 0) {
          // v_a^H v_b = conj(inv_a) inv_b (x_a^H x_b)(rows > b) + conj(inv_a x_a[b]) * 1
          const double2 ia = s_inv[a], ib = s_inv[b];
          const double2 t = cmul(cmulc(ia, ib), make_double2(r, im));
          const double2 h = cmul(ia, Pn[b + (size_t)ld * a]);
          sS[a][b] = make_double2(t.x + h.x, t.y - h.y);
        }
      }
  }
  for (int i = tid; i < PB * PB; i += PT) sT[i / PB][i % PB] = make_double2(0.0, 0.0);
  __syncthreads();
  // zlarft (forward, columnwise): T(j,j) = tau_j ; T(0:j, j) = -tau_j * T(0:j,0:j) * S(0:j, j)
  if (warp == 0) {
    for (int j = 0; j < pw; ++j) {
      const double2 tau = s_tau[j];
      if (lane < j) {
        double2 acc = make_double2(0.0, 0.0);
        for (int l = lane; l < j; ++l) { const double2 t = cmul(sT[lane][l], sS[l][j]); acc.x += t.x; acc.y += t.y; }
        const double2 t = cmul(tau, acc);
        sT[lane][j] = make_double2(-t.x, -t.y);
      }
      if (lane == j) sT[j][j] = tau;
      __syncwarp();
    }
  }
  __syncthreads();
  double2* Tout = P.T + (size_t)(k & 1) * PB * PB;          // V and T are double-buffered: the trailing update of panel k
  double2* Vout = P.V + (size_t)(k & 1) * P.M * PB;         // may still be running while panel k+1 is factored
  for (int i = tid; i < PB * PB; i += PT) Tout[i] = sT[i % PB][i / PB];   // column-major PB x PB
  // clean reflector block for the update kernel: V (mk x PB, ld = P.M), unit diagonal, zeros above and beyond pw;
  // and the panel itself back to global memory
  {
    const int total = mk * PB;
    for (int e = tid; e < total; e += PT) {
      const int c = e / mk, i = e - c * mk;
      double2 v = make_double2(0.0, 0.0);
      if (c < pw) {
        const double2 x = Pn[i + (size_t)ld * c];
        if (in_smem) gp[i + (size_t)P.ldy * c] = x;
        if (i == c) v = make_double2(1.0, 0.0);
        else if (i > c) v = cmul(x, s_inv[c]);
      }
      Vout[i + (size_t)P.M * c] = v;
    }
  }
}

// C <- C - V * (T^H * (V^H C)) on a slab of SLAB trailing columns
// trailing columns [t0 + lo, t0 + hi) only (t0 = first column after the panel): the look-ahead schedule updates the next
// panel's columns (lo = 0, hi = PB) ahead of the rest (lo = PB, hi = inf)
__global__ void __launch_bounds__(UT, 4) qr_update_kernel(const QrProblem* __restrict__ probs, int k, int lo, int hi) {
  const QrProblem P = probs[blockIdx.y];
  const int c0 = k * PB;
  if (c0 >= P.N) return;
  const int pw = min(PB, P.N - c0);
  const int t0 = c0 + pw;
  const int col0 = t0 + lo + blockIdx.x * SLAB;
  const int cend = (int)min((long)P.N, (long)t0 + hi);
  if (col0 >= cend) return;
  const int ncol = min(SLAB, cend - col0);
  const int r0 = c0, mk = P.M - r0;
  const double2* __restrict__ V = P.V + (size_t)(k & 1) * P.M * PB;
  const double2* __restrict__ Tk = P.T + (size_t)(k & 1) * PB * PB;
  const int ldv = P.M, ldy = P.ldy;
  double2* C = P.Y + r0 + (size_t)ldy * col0;

  __shared__ double2 sW[4][PB][SLAB + 1];
  __shared__ double2 sW2[PB][SLAB + 1];
  __shared__ double2 sTm[PB][PB + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lk = lane & 3;
  for (int i = tid; i < PB * PB; i += UT) sTm[i % PB][i / PB] = Tk[i];

  // ---------------- phase 1: W = V^H C  (PB x SLAB), K = mk rows split over the warps
  {
    double wre[2][4][2], wim[2][4][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n) { wre[m][n][0] = wre[m][n][1] = wim[m][n][0] = wim[m][n][1] = 0.0; }
    const int nch = (mk + 3) >> 2;
    for (int ch = warp; ch < nch; ch += 4) {
      const int row = 4 * ch + lk;
      const bool rok = row < mk;
      double2 a[2], b[4];
#pragma unroll
      for (int m = 0; m < 2; ++m) a[m] = rok ? V[row + (size_t)ldv * (8 * m + lr)] : make_double2(0.0, 0.0);
#pragma unroll
      for (int n = 0; n < 4; ++n) b[n] = (rok && 8 * n + lr < ncol) ? C[row + (size_t)ldy * (8 * n + lr)] : make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const double ar = a[m].x, ai = a[m].y, nai = -a[m].y;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          dmma884(wre[m][n][0], wre[m][n][1], ar, b[n].x);
          dmma884(wre[m][n][0], wre[m][n][1], ai, b[n].y);
          dmma884(wim[m][n][0], wim[m][n][1], ar, b[n].y);
          dmma884(wim[m][n][0], wim[m][n][1], nai, b[n].x);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) sW[warp][8 * m + lr][8 * n + 2 * lk + e] = make_double2(wre[m][n][e], wim[m][n][e]);
  }
  __syncthreads();
  // ---------------- phase 2: W2 = T^H (sum of the four partial W)
  for (int i = tid; i < PB * SLAB; i += UT) {
    const int r = i / SLAB, c = i % SLAB;
    const double2 a = sW[0][r][c], b = sW[1][r][c], cc = sW[2][r][c], d = sW[3][r][c];
    sW[0][r][c] = make_double2(a.x + b.x + cc.x + d.x, a.y + b.y + cc.y + d.y);
  }
  __syncthreads();
  for (int i = tid; i < PB * SLAB; i += UT) {
    const int r = i / SLAB, c = i % SLAB;
    double2 acc = make_double2(0.0, 0.0);
    for (int l = 0; l <= r; ++l) { const double2 t = cmulc(sTm[l][r], sW[0][l][c]); acc.x += t.x; acc.y += t.y; }
    sW2[r][c] = acc;
  }
  __syncthreads();
  // ---------------- phase 3: C -= V W2
  {
    const int nch = (mk + 7) >> 3;
    for (int ch = warp; ch < nch; ch += 4) {
      const int row = 8 * ch + lr;
      const bool rok = row < mk;
      double2 a[4];
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) a[kc] = rok ? V[row + (size_t)ldv * (4 * kc + lk)] : make_double2(0.0, 0.0);
      double cre[4][2], cim[4][2];
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * n + 2 * lk + e;
          const double2 v = (rok && col < ncol) ? C[row + (size_t)ldy * col] : make_double2(0.0, 0.0);
          cre[n][e] = v.x; cim[n][e] = v.y;
        }
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        const double nar = -a[kc].x, ai = a[kc].y, nai = -a[kc].y;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const double2 b = sW2[4 * kc + lk][8 * n + lr];
          dmma884(cre[n][0], cre[n][1], nar, b.x);
          dmma884(cre[n][0], cre[n][1], ai, b.y);
          dmma884(cim[n][0], cim[n][1], nar, b.y);
          dmma884(cim[n][0], cim[n][1], nai, b.x);
        }
      }
      if (rok) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 8 * n + 2 * lk + e;
            if (col < ncol) C[row + (size_t)ldy * col] = make_double2(cre[n][e], cim[n][e]);
          }
      }
    }
  }
}

// G = R^H : G[i + N j] = conj(Y[j + ldy i]) for j <= i, else 0
__global__ void __launch_bounds__(256) qr_rh_kernel(const QrProblem* __restrict__ probs) {
  const QrProblem P = probs[blockIdx.z];
  const int N = P.N;
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  if (i0 >= N || j0 >= N) return;
  __shared__ double2 tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // read Y[j0 + tx][i0 + r]  (rows j contiguous)
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + tx, i = i0 + r;
    double2 v = make_double2(0.0, 0.0);
    if (i < N && j < N && j <= i) { v = P.Y[j + (size_t)P.ldy * i]; v.y = -v.y; }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + tx, j = j0 + r;
    if (i < N && j < N) P.G[i + (size_t)N * j] = tile[tx][r];
  }
}

}  // namespace

void launch_qr(const QrProblem* d_probs, int batch, int max_m, int max_n, cudaStream_t s) {
  if (batch <= 0 || max_n <= 0) return;
  // the panel is staged in shared memory whenever it fits (the kernel applies the same test per matrix)
  const size_t want = (size_t)max_m * PB * sizeof(double2);
  const int smem = (int)(want < (size_t)PANEL_SMEM_CAP ? want : (size_t)PANEL_SMEM_CAP);
  static bool attr_set_dev[64] = {};   // per device context (several devices in one process: site-sharded handles)
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr_set = attr_set_dev[dev & 63];
  if (!attr_set) {
    cudaFuncSetAttribute(qr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_CAP);
    attr_set = true;
  }
  const int npanels = (max_n + PB - 1) / PB;
  for (int k = 0; k < npanels; ++k) {
    qr_panel_kernel<<<batch, PT, smem, s>>>(d_probs, k);
    const int ntr = max_n - (k + 1) * PB;
    if (ntr <= 0) break;
    dim3 grid((ntr + SLAB - 1) / SLAB, batch);
    qr_update_kernel<<<grid, UT, 0, s>>>(d_probs, k, 0, 1 << 30);
  }
  dim3 g2((max_n + 31) / 32, (max_n + 31) / 32, batch);
  qr_rh_kernel<<<g2, 256, 0, s>>>(d_probs);
}
int qr_launch_count(int max_n) {
  const int npanels = (max_n + PB - 1) / PB;
  return 2 * npanels;   // panels + trailing updates (the last panel has none) + the R^H transpose
}

}  // namespace mpsb200
