// HBM-bound site kernels of the MPS engine: in-place single-qubit gate (ExaTnMpsVisitor.cpp:1185-1292),
// environment trace for the <Z>/<ZZ> sweeps, fills and the device RNG used by the micro-benchmarks.
#include "kernels.h"

namespace mpsb200 {
namespace {

// new[a,b,c] = sum_i m[b][i] old[a,i,c]; site is column-major (dl, 2, dr): the two physical slices of a
// right-bond column c are the contiguous runs [2 dl c, 2 dl c + dl) and [2 dl c + dl, 2 dl (c+1)).
// One thread per (a,c): two coalesced 16-byte loads, two stores; 64 bytes of traffic per 32 flop.
__global__ void __launch_bounds__(256) gate1q_kernel(const Gate1qProblem* __restrict__ probs) {
  const Gate1qProblem* P = probs + blockIdx.y;
  const int dl = P->dl;
  const long total = (long)dl * P->dr;
  const double2 m00 = P->m[0], m01 = P->m[1], m10 = P->m[2], m11 = P->m[3];
  double2* __restrict__ t = P->site;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long c = e / dl;
    const long a = e - c * dl;
    const long i0 = a + 2 * dl * c, i1 = i0 + dl;
    const double2 x0 = t[i0], x1 = t[i1];
    double2 y0, y1;
    y0.x = m00.x * x0.x - m00.y * x0.y + m01.x * x1.x - m01.y * x1.y;
    y0.y = m00.x * x0.y + m00.y * x0.x + m01.x * x1.y + m01.y * x1.x;
    y1.x = m10.x * x0.x - m10.y * x0.y + m11.x * x1.x - m11.y * x1.y;
    y1.y = m10.x * x0.y + m10.y * x0.x + m11.x * x1.y + m11.y * x1.x;
    t[i0] = y0;
    t[i1] = y1;
  }
}

__global__ void __launch_bounds__(256) trace_pair_kernel(const double2* __restrict__ E, const double2* __restrict__ R, int n,
                                                         double2* __restrict__ out) {
  // single CTA: sum_{i,j} E[i + n j] * R[j + n i]
  double re = 0.0, im = 0.0;
  const long total = (long)n * n;
  for (long e = threadIdx.x; e < total; e += blockDim.x) {
    const long j = e / n, i = e - j * n;
    const double2 a = E[e], b = R[j + (long)n * i];
    re += a.x * b.x - a.y * b.y;
    im += a.x * b.y + a.y * b.x;
  }
  __shared__ double sr[256], si[256];
  sr[threadIdx.x] = re; si[threadIdx.x] = im;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = make_double2(sr[0], si[0]);
}

// out = sum_{a,p,c} w_p conj(F[a,p,c]) H[a,p,c]  over a site-shaped pair (dl, 2, dr); single CTA, fixed reduction order.
// With F_p = L S_p (left sweep) and H_p = S_p R (right sweep) and Hermitian L this is <psi| diag(w0, w1)_k |psi>.
__global__ void __launch_bounds__(1024) site_dot_kernel(const double2* __restrict__ F, const double2* __restrict__ H, int dl, int dr,
                                                        double w0, double w1, double2* __restrict__ out) {
  double re = 0.0, im = 0.0;
  const long total = 2L * dl * dr;
  for (long e = threadIdx.x; e < total; e += 1024) {
    const int p = (int)((e / dl) & 1);
    const double w = p ? w1 : w0;
    const double2 a = F[e], b = H[e];
    re += w * (a.x * b.x + a.y * b.y);
    im += w * (a.x * b.y - a.y * b.x);
  }
  __shared__ double sr[1024], si[1024];
  sr[threadIdx.x] = re; si[threadIdx.x] = im;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = make_double2(sr[0], si[0]);
}

// Batched, multi-CTA version of the two reductions above for the observables: problem b is either a weighted site dot
// (mode 0: sum_e w_p(e) conj(F[e]) H[e] over a site-shaped pair) or an environment trace (mode 1: sum_{i,j} F[i + n j] H[j + n i]).
// Stage 1: CTA (x, b) reduces a strided share into partial[b][x]; stage 2 sums the partials of a problem in a fixed order.
__global__ void __launch_bounds__(256) dot_batch_partial_kernel(const DotProblem* __restrict__ probs, double2* __restrict__ partial) {
  const DotProblem P = probs[blockIdx.y];
  double re = 0.0, im = 0.0;
  if (P.mode == 0) {
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < P.total; e += (long)gridDim.x * 256) {
      const int p = (int)((e / P.n) & 1);
      const double w = p ? P.w1 : P.w0;
      const double2 a = P.F[e], b = P.H[e];
      re += w * (a.x * b.x + a.y * b.y);
      im += w * (a.x * b.y - a.y * b.x);
    }
  } else {
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < P.total; e += (long)gridDim.x * 256) {
      const long j = e / P.n, i = e - j * P.n;
      const double2 a = P.F[e], b = P.H[j + (long)P.n * i];
      re += a.x * b.x - a.y * b.y;
      im += a.x * b.y + a.y * b.x;
    }
  }
  __shared__ double sr[256], si[256];
  sr[threadIdx.x] = re; si[threadIdx.x] = im;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = make_double2(sr[0], si[0]);
}
__global__ void dot_batch_final_kernel(const double2* __restrict__ partial, int per_problem, int nprob, double2* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nprob) return;
  double re = 0.0, im = 0.0;
  for (int i = 0; i < per_problem; ++i) { const double2 v = partial[(size_t)b * per_problem + i]; re += v.x; im += v.y; }
  out[b] = make_double2(re, im);
}

__global__ void fill_kernel(double2* p, long n, double2 v) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void randn_kernel(double2* p, long n, uint64_t seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const uint64_t a = splitmix(seed + 2 * (uint64_t)i), b = splitmix(seed + 2 * (uint64_t)i + 1);
    const double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (b >> 11) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    p[i] = make_double2(r * cs, r * sn);
  }
}
}  // namespace

void launch_gate1q(const Gate1qProblem* d_probs, int batch, long max_elems, cudaStream_t s) {
  if (batch <= 0 || max_elems <= 0) return;
  long bx = (max_elems + 255) / 256;
  if (bx > 148L * 8) bx = 148L * 8;   // grid-stride; multiples of the SM count once saturated
  dim3 grid((unsigned)bx, batch);
  gate1q_kernel<<<grid, 256, 0, s>>>(d_probs);
}
void launch_trace_pair(const double2* E, const double2* R, int n, double2* out, cudaStream_t s) {
  trace_pair_kernel<<<1, 256, 0, s>>>(E, R, n, out);
}
void launch_site_dot(const double2* F, const double2* H, int dl, int dr, double w0, double w1, double2* out, cudaStream_t s) {
  site_dot_kernel<<<1, 1024, 0, s>>>(F, H, dl, dr, w0, w1, out);
}
void launch_dot_batch(const DotProblem* d_probs, int nprob, int ctas_per_problem, double2* d_partial, double2* d_out, cudaStream_t s) {
  if (nprob <= 0) return;
  if (ctas_per_problem == 1) {   // one CTA per problem: its "partial" is the result
    dot_batch_partial_kernel<<<dim3(1, (unsigned)nprob), 256, 0, s>>>(d_probs, d_out);
    return;
  }
  dot_batch_partial_kernel<<<dim3((unsigned)ctas_per_problem, (unsigned)nprob), 256, 0, s>>>(d_probs, d_partial);
  dot_batch_final_kernel<<<(nprob + 127) / 128, 128, 0, s>>>(d_partial, ctas_per_problem, nprob, d_out);
}
void launch_fill(double2* p, long n, double2 v, cudaStream_t s) {
  if (n <= 0) return;
  long bx = (n + 255) / 256;
  if (bx > 148L * 8) bx = 148L * 8;
  fill_kernel<<<(unsigned)bx, 256, 0, s>>>(p, n, v);
}
void launch_randn(double2* p, long n, uint64_t seed, cudaStream_t s) {
  if (n <= 0) return;
  long bx = (n + 255) / 256;
  if (bx > 148L * 8) bx = 148L * 8;
  randn_kernel<<<(unsigned)bx, 256, 0, s>>>(p, n, seed);
}

}  // namespace mpsb200
