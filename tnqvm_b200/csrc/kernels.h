// Device-side problem descriptors and kernel launchers of the B200 MPS engine (sm_100a).
// All tensors are complex128 (double2), column-major.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpsb200 {

// ---------------------------------------------------------------------------------------------
// Batched complex-FP64 GEMM on DMMA (mma.sync m8n8k4 f64) with cp.async.bulk (TMA engine) staging.
//   C = alpha * rowscale .* (opA(A) * opB(B)) .* colscale        (plain mode)
//   theta mode: the four products A_p * B_q (p,q in {0,1}) of an MPS site pair are accumulated in
//   one CTA and mixed by the 4x4 gate in the epilogue (ExaTnMpsVisitor.cpp:1394-1549 in one kernel).
struct GemmProblem {
  const double2* A;
  const double2* B;
  double2* C;
  double2* C2;              // optional second copy of the output (the SVD works in place on it)
  const int* a_gather;      // AK layout: source column of logical row i (nullptr = identity)
  const int* b_gather;      // BK layout: source column of logical column j (nullptr = stride/off)
  const double* row_scale;  // optional, length M
  const double* col_scale;  // optional, length N
  int M, N, K;              // theta mode: M = chi_L, N = chi_R (per physical sub-block), K = chi
  int lda, ldb, ldc;
  int b_col_stride, b_col_off;
  int mode;                 // 0 plain, 1 theta
  int conjT_out;            // store C^H instead of C
  double alpha;
  double2 gate[16];         // theta mode: row-major 4x4 in (p_lo,p_hi) order
};
// layout: 0 = "NN" (A MxK col-major, B KxN col-major), 1 = "CN" (A given as K x M, used conj-transposed),
//         2 = "NC" (B given as N x K, used conj-transposed)
void launch_gemm(const GemmProblem* d_probs, int batch, int max_tiles, int layout, cudaStream_t s);
void launch_gemm1(const GemmProblem& p, int layout, cudaStream_t s);   // one problem, descriptor by value
int gemm_tiles(int M, int N, int mode);
void gemm_set_small_path(int on);   // A/B switch for the 32x32-tile single-problem kernel (default on)

// ---------------------------------------------------------------------------------------------
// Batched blocked Householder QR, R factor only (qr_dmma.cu): Y (M x N, M >= N) is overwritten (R in its upper
// triangle), G (N x N, ld = N) receives R^H.  V (2 x M x QR_PB) and T (2 x QR_PB x QR_PB) are per-matrix scratch
// (double-buffered over consecutive panels).
constexpr int QR_PB = 16;
struct QrProblem {
  double2* Y;
  double2* V;
  double2* T;
  double2* G;
  int M, N, ldy;
};
void launch_qr(const QrProblem* d_probs, int batch, int max_m, int max_n, cudaStream_t s);
int qr_launch_count(int max_n);

// ---------------------------------------------------------------------------------------------
// Batched one-sided Jacobi SVD (block Hestenes): G (M x N, M >= N) -> G V with orthogonal columns.
struct JacobiProblem {
  double2* G;
  int M, N, ldg;
  int nb;    // column blocks of 8 (ceil(N/8))
  int nbe;   // nb rounded up to even (>= 2) ; 1 when nb == 1
  double2* wd;   // nb travelling 8x8 Gram blocks (row-major, 64 complex each): W_BB of every column block
  int* ver;      // nb block versions (rotations seen), zeroed per SVD
  int2* rec;     // nbe x nbe clean-pair memo: versions (+1) of (A < B) at which their cross pairs were last found clean; zeroed per SVD
};
// one whole sweep (nsteps tournament steps) in one persistent launch; d_progress: per matrix `progress_stride` ints (>= nbe), zeroed once
// per SVD; base = steps completed by the previous sweeps; *d_counter zeroed (one counter per launch).  d_active: [0] = matrices
// still rotating (read on the device: the host's `batch` is only an upper bound used to size the grid), [1..] their indices.
void launch_jacobi_sweep(const JacobiProblem* d_probs, int batch, int max_pairs, int nsteps, int base, double tol2, double dead2,
                         const double* d_fro2, int* d_dirty, const int* d_done, int* d_progress, int progress_stride, int* d_counter,
                         int* d_fault /* += 1 if a dependency wait timed out */, int grid_ctas, int warps_per_task /* 4 or 8 */,
                         const int* d_active, cudaStream_t s);
// The same sweep with shared-memory-resident pair tasks split by rows over thread-block clusters (jacobi_cluster.cu).  Progress flags:
// jacobi_cluster_progress_ints_per_block() ints per 8-column block (one per cluster rank); no task counter (static assignment).
int jacobi_cluster_progress_ints_per_block();
bool jacobi_cluster_shape(int max_m, int* cs, int* rpc_cap);          // cluster size / rows per CTA for matrices of up to max_m rows
int jacobi_cluster_max_clusters(int cs, int rpc_cap);                 // co-resident clusters on the current device
int launch_jacobi_cluster_sweep(const JacobiProblem* d_probs, int batch, int max_pairs, int nsteps, int base, double tol2, double dead2,
                                const double* d_fro2, int* d_dirty, const int* d_done, int* d_progress, int progress_stride, int* d_fault,
                                const int* d_active, int cs, int rpc_cap, int max_clusters, cudaStream_t s);   // CTAs launched, -1 on error
double jacobi_cluster_dmma_flops();
void jacobi_cluster_set_debug(int mode);   // timing experiments only
void jacobi_cluster_print_phase_timing();
void jacobi_set_debug_mode(int mode);   // timing experiments only
void jacobi_print_phase_timing();
double jacobi_dmma_flops();   // process-wide count of the FP64 flops the Jacobi pair tasks executed (Gram on DMMA + scaled rotations)
void launch_fro2(const JacobiProblem* d_probs, int batch, double* d_fro2, cudaStream_t s);   // d_fro2 pre-zeroed
// after a sweep: done[m] |= !dirty[m]; dirty[m] = 0; *remaining = #not done
void launch_jacobi_check(int batch, int* d_dirty, int* d_done, int* d_remaining, int* d_active /* may be null */, cudaStream_t s);

// column norms, descending sort, truncation rule (ExaTnMpsVisitor.cpp:2434-2445) and write-back scales
struct TruncProblem {
  const double2* G;
  int M, N, ldg;
  int tall;          // 1: G = theta V (columns ~ U S) ; 0: G = theta^H V' (columns ~ V S)
  double* sig2;      // scratch N
  double* sigma;     // out: sorted singular values (N)
  int* perm;         // out: sorted -> column
  double* scaleP;    // out: per kept column scale for the factor taken from G
  double* scaleO;    // out: per kept column scale for the factor obtained by GEMM
  int* keep;         // out
  double* weights;   // out[2]: total weight, kept weight
};
void launch_trunc(const TruncProblem* d_probs, int batch, double cutoff, int cutoff_on_sqrt, int max_bond, int gauge,
                  int renorm, double null_tol, cudaStream_t s);

// gather/scale copy of kept columns: out[i + ldo*k] = G[i + ldg*perm[k]] * scale[k]      (conjT = 0)
//                                   out[k + ldo*i] = conj(G[i + ldg*perm[k]]) * scale[k] (conjT = 1)
struct GatherProblem {
  const double2* G;
  double2* out;
  const int* perm;
  const double* scale;
  int M, keep, ldg, ldo, conjT;   // keep: kept columns (known on the host after the truncation read-back)
};
void launch_gather(const GatherProblem* d_probs, int batch, int max_rows, int max_keep, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// single-qubit gate, in place on a site tensor (ExaTnMpsVisitor.cpp:1185-1292)
struct Gate1qProblem {
  double2* site;
  int dl, dr;
  double2 m[4];
};
void launch_gate1q(const Gate1qProblem* d_probs, int batch, long max_elems, cudaStream_t s);

// out[0] = sum_{i,j} E[i + n*j] * R[j + n*i]   (trace(E R)), complex
void launch_trace_pair(const double2* E, const double2* R, int n, double2* out, cudaStream_t s);
// out[0] = sum_{a,p,c} w_p conj(F[a,p,c]) H[a,p,c]  (site-shaped arrays (dl,2,dr), column-major)
void launch_site_dot(const double2* F, const double2* H, int dl, int dr, double w0, double w1, double2* out, cudaStream_t s);
// batched multi-CTA reductions of the observables (two launches, fixed summation order)
struct DotProblem {
  const double2* F;
  const double2* H;
  long total;      // elements of F
  int n;           // mode 0: dl of the site-shaped pair (physical index = (e / dl) & 1); mode 1: order of the square environments
  int mode;        // 0: sum w_p conj(F) H ; 1: sum_{i,j} F[i + n j] H[j + n i]
  double w0, w1;
};
void launch_dot_batch(const DotProblem* d_probs, int nprob, int ctas_per_problem, double2* d_partial /* nprob x ctas */, double2* d_out, cudaStream_t s);
void launch_fill(double2* p, long n, double2 v, cudaStream_t s);
void launch_randn(double2* p, long n, uint64_t seed, cudaStream_t s);

}  // namespace mpsb200
