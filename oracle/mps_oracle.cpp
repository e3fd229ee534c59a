// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the `exatn-mps` visitor algorithm of ORNL-QCI/tnqvm
// (tnqvm/visitors/exatn-mps/ExaTnMpsVisitor.cpp) in plain C++ over host
// BLAS/LAPACK (scipy's bundled OpenBLAS, resolved with dlopen at run time).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// The floating-point work of the reference lives in ExaTN/TAL-SH
// (github.com/ornl-qci/exatn, unpinned: the reference CI clones HEAD,
// .github/workflows/build.yml:33), which is absent from /root/reference.  Its
// semantics are restated from the call sites:
//   contractTensorsSync      ExaTnMpsVisitor.cpp:1253,1439,1520  -> zgemm / loops
//   decomposeTensorSVDLRSync ExaTnMpsVisitor.cpp:1623            -> zgesvd|zgesdd, sqrt(S) into both factors
//   computePartialNormsSync  ExaTnMpsVisitor.cpp:2425,2428       -> per-bond-slice sum of squares (comment :2421-2423)
//   extractTensorSliceSync   ExaTnMpsVisitor.cpp:2482,2485       -> leading slices of the bond
//   evaluateSync(ket)        ExaTnMpsVisitor.cpp:595             -> chain of zgemm, qubit 0 = LSB
// Parity pinning: see tests/test_oracle_golden.py (reference gtest known answers,
// SURVEY.md section 8c) and oracle/dense_ref.cpp (the reference's own dense simulator).
// Truncated-run behaviour is NOT pinned by any reference test ("parity unpinned"
// for max-bond-dim / svd-cutoff active).
#include <algorithm>
#include <array>
#include <cfloat>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <random>
#include <string>
#include <vector>

typedef std::complex<double> cplx;

// ---------------------------------------------------------------- BLAS/LAPACK
typedef void (*zgemm_t)(const char*, const char*, const int*, const int*, const int*, const cplx*, const cplx*, const int*,
                        const cplx*, const int*, const cplx*, cplx*, const int*);
typedef void (*zgesvd_t)(const char*, const char*, const int*, const int*, cplx*, const int*, double*, cplx*, const int*,
                         cplx*, const int*, cplx*, const int*, double*, int*);
typedef void (*zgesdd_t)(const char*, const int*, const int*, cplx*, const int*, double*, cplx*, const int*, cplx*,
                         const int*, cplx*, const int*, double*, int*, int*);
typedef void (*setthr_t)(int);
static zgemm_t p_zgemm = nullptr;
static zgesvd_t p_zgesvd = nullptr;
static zgesdd_t p_zgesdd = nullptr;
static setthr_t p_setthr = nullptr;

extern "C" int oracle_init_blas(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "oracle: dlopen(%s) failed: %s\n", path, dlerror()); return 1; }
  p_zgemm = (zgemm_t)dlsym(h, "scipy_zgemm_");
  p_zgesvd = (zgesvd_t)dlsym(h, "scipy_zgesvd_");
  p_zgesdd = (zgesdd_t)dlsym(h, "scipy_zgesdd_");
  p_setthr = (setthr_t)dlsym(h, "scipy_openblas_set_num_threads");
  if (!p_zgemm) p_zgemm = (zgemm_t)dlsym(h, "zgemm_");
  if (!p_zgesvd) p_zgesvd = (zgesvd_t)dlsym(h, "zgesvd_");
  if (!p_zgesdd) p_zgesdd = (zgesdd_t)dlsym(h, "zgesdd_");
  if (!p_setthr) p_setthr = (setthr_t)dlsym(h, "openblas_set_num_threads");
  return (p_zgemm && p_zgesvd && p_zgesdd) ? 0 : 2;
}
extern "C" void oracle_set_threads(int n) { if (p_setthr) p_setthr(n); }

// C(MxN) = A(MxK) * B(KxN), all column-major.
static void gemm(int M, int N, int K, const cplx* A, int lda, const cplx* B, int ldb, cplx* C, int ldc, char ta = 'N') {
  const cplx one(1, 0), zero(0, 0);
  if (M == 0 || N == 0) return;
  if (p_zgemm && (long)M * N * K > 512) {
    p_zgemm(&ta, "N", &M, &N, &K, &one, A, &lda, B, &ldb, &zero, C, &ldc);
    return;
  }
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < M; ++i) {
      cplx s = 0;
      for (int k = 0; k < K; ++k) {
        cplx a = (ta == 'N') ? A[i + (size_t)lda * k] : std::conj(A[k + (size_t)lda * i]);
        s += a * B[k + (size_t)ldb * j];
      }
      C[i + (size_t)ldc * j] = s;
    }
}

// ---------------------------------------------------------------- gate matrices
// Restates tnqvm/base/Gates.hpp:132-334 (row = output index).  Unknown names -> identity
// (ExatnUtils.cpp:112).  Returns the matrix dimension (2 or 4), row-major into out.
extern "C" int oracle_gate_matrix(const char* name_c, const double* params, int nparams, double* out_ri) {
  std::string name(name_c);
  const cplx I(0, 1);
  auto p = [&](int i) { return i < nparams ? params[i] : 0.0; };
  std::vector<cplx> m;
  int dim = 2;
  if (name == "CX") name = "CNOT";
  if (name == "H") m = {M_SQRT1_2, M_SQRT1_2, M_SQRT1_2, -M_SQRT1_2};
  else if (name == "X") m = {0, 1, 1, 0};
  else if (name == "Y") m = {0, -I, I, 0};
  else if (name == "Z") m = {1, 0, 0, -1};
  else if (name == "Rx") m = {std::cos(0.5 * p(0)), -I * std::sin(0.5 * p(0)), -I * std::sin(0.5 * p(0)), std::cos(0.5 * p(0))};
  else if (name == "Ry") m = {std::cos(0.5 * p(0)), -std::sin(0.5 * p(0)), std::sin(0.5 * p(0)), std::cos(0.5 * p(0))};
  else if (name == "Rz") m = {std::exp(cplx(0, -0.5 * p(0))), 0, 0, std::exp(cplx(0, 0.5 * p(0)))};
  else if (name == "T") m = {1, 0, 0, std::exp(cplx(0, M_PI_4))};
  else if (name == "Tdg") m = {1, 0, 0, std::exp(cplx(0, -M_PI_4))};
  else if (name == "U")
    m = {std::cos(p(0) / 2.0), -std::exp(cplx(0, p(2))) * std::sin(p(0) / 2.0), std::exp(cplx(0, p(1))) * std::sin(p(0) / 2.0),
         std::exp(cplx(0, p(1) + p(2))) * std::cos(p(0) / 2.0)};
  else if (name == "CNOT") { dim = 4; m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0}; }
  else if (name == "CZ") { dim = 4; m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1}; }
  else if (name == "CY") { dim = 4; m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, -I, 0, 0, I, 0}; }
  else if (name == "CH") { dim = 4; m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, M_SQRT1_2, M_SQRT1_2, 0, 0, M_SQRT1_2, -M_SQRT1_2}; }
  else if (name == "CRZ") { dim = 4; m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, std::exp(cplx(0, -0.5 * p(0))), 0, 0, 0, 0, std::exp(cplx(0, 0.5 * p(0)))}; }
  else if (name == "CPhase") { dim = 4; m = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, std::exp(cplx(0, p(0)))}; }
  else if (name == "Swap") { dim = 4; m = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1}; }
  else if (name == "iSwap") { dim = 4; m = {1, 0, 0, 0, 0, 0, I, 0, 0, I, 0, 0, 0, 0, 0, 1}; }
  else if (name == "fSim") {
    dim = 4;
    m = {1, 0, 0, 0, 0, std::cos(p(0)), cplx(0, -std::sin(p(0))), 0, 0, cplx(0, -std::sin(p(0))), std::cos(p(0)), 0, 0, 0, 0,
         std::exp(cplx(0, -p(1)))};
  } else m = {1, 0, 0, 1};  // I and anything unknown
  for (size_t i = 0; i < m.size(); ++i) { out_ri[2 * i] = m[i].real(); out_ri[2 * i + 1] = m[i].imag(); }
  return dim;
}

// ---------------------------------------------------------------- the MPS
struct Site {
  int dl = 1, dr = 1;     // left / right bond; storage column-major (dl, 2, dr)
  std::vector<cplx> t;    // first site == (1,2,dr) == reference's (phys,right); last == (dl,2,1) == (left,phys)
};
struct Oracle {
  int n = 0;
  std::vector<Site> s;
  double svd_cutoff = DBL_MIN;   // ExaTnMpsVisitor.cpp:257-263
  int max_bond = INT_MAX - 1;    // ExaTnMpsVisitor.cpp:265-271
  int cutoff_on_sqrt = 0;        // 0: partial norm = sum of squares (= sigma_k); 1: 2-norm (= sqrt(sigma_k))
  int use_gesdd = 0;
  int gauge = 0;                 // 0 reference (U sqrtS | sqrtS Vh); 1 (U | S Vh); 2 (U S | Vh)
  int renorm = 0;                // never in the reference
  double null_tol = 0;           // > 0: singular values <= null_tol * ||theta||_F are rounding noise of an exact zero and are set to 0.0 before the
                                 // reference's cut rule runs (NOT in the reference: it mirrors the engine's documented deviation)
  std::vector<std::vector<double>> sv;   // last retained singular values per bond
  double discarded = 0;          // accumulated discarded weight (not in reference; diagnostic)
  double log_fidelity = 0;       // sum of log(1 - w) over truncations: fidelity estimate prod(1 - w) (diagnostic, config 5)
  double t_gemm = 0, t_svd = 0, t_trunc = 0;
  std::mt19937_64 rng;           // RandomEngine.hpp:43 (process-global there; per-handle here)
  std::vector<int> measure;      // visit(Measure) order, ExaTnMpsVisitor.cpp:991-994
};

// ExaTnMpsVisitor.cpp:281-326 : |0...0>, all bonds 1
extern "C" Oracle* oracle_create(int n, int max_bond, double svd_cutoff, int cutoff_on_sqrt, int use_gesdd, int gauge) {
  Oracle* o = new Oracle;
  o->n = n;
  o->s.resize(n);
  for (auto& s : o->s) { s.dl = s.dr = 1; s.t = {cplx(1, 0), cplx(0, 0)}; }
  if (max_bond > 0) o->max_bond = max_bond;
  if (svd_cutoff >= 0) o->svd_cutoff = svd_cutoff;
  o->cutoff_on_sqrt = cutoff_on_sqrt;
  o->use_gesdd = use_gesdd;
  o->gauge = gauge;
  o->sv.assign(n > 0 ? n - 1 : 0, std::vector<double>{1.0});
  std::random_device rd;
  o->rng.seed(rd());
  return o;
}
extern "C" void oracle_destroy(Oracle* o) { delete o; }
extern "C" void oracle_seed(Oracle* o, uint64_t seed) { o->rng.seed(seed); }   // TNQVM.hpp:114-117
extern "C" void oracle_set_renorm(Oracle* o, int r) { o->renorm = r; }
extern "C" void oracle_set_null_tol(Oracle* o, double t) { o->null_tol = t; }

// applyGate 1q branch, ExaTnMpsVisitor.cpp:1185-1292: new[b] = sum_i M[b][i] old[i] on the physical leg.
extern "C" void oracle_apply_1q(Oracle* o, int q, const double* m_ri) {
  const cplx* m = reinterpret_cast<const cplx*>(m_ri);
  Site& s = o->s[q];
  for (int c = 0; c < s.dr; ++c)
    for (int a = 0; a < s.dl; ++a) {
      cplx& x0 = s.t[a + (size_t)s.dl * (0 + 2 * c)];
      cplx& x1 = s.t[a + (size_t)s.dl * (1 + 2 * c)];
      cplx y0 = m[0] * x0 + m[1] * x1, y1 = m[2] * x0 + m[3] * x1;
      x0 = y0; x1 = y1;
    }
}

static double now_s() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// applyTwoQubitGate, ExaTnMpsVisitor.cpp:1387-1731 + truncateSvdTensors :2366-2536.
// m = row-major 4x4; matrix index = 2*bit(q0) + bit(q1) whichever site is left (:1492-1499).
extern "C" int oracle_apply_2q(Oracle* o, int q0, int q1, const double* m_ri) {
  if (std::abs(q0 - q1) != 1) return 1;   // assert at :1398
  const cplx* m = reinterpret_cast<const cplx*>(m_ri);
  const int lo = std::min(q0, q1);
  Site& A = o->s[lo];
  Site& B = o->s[lo + 1];
  const int cl = A.dl, ch = A.dr, cr = B.dr;
  const int M = 2 * cl, N = 2 * cr;
  double t0 = now_s();
  // step 1 merge (:1394-1440): D(a,p,q,c) = sum_k A(a,p,k) B(k,q,c) ; as (2cl x ch)*(ch x 2cr)
  std::vector<cplx> D((size_t)M * N), T((size_t)M * N);
  gemm(M, N, ch, A.t.data(), M, B.t.data(), ch, D.data(), M);
  // step 2 gate (:1442-1549)
  const bool q0_is_lo = (q0 == lo);
  for (int c = 0; c < cr; ++c)
    for (int a = 0; a < cl; ++a) {
      cplx in[4], out[4];
      for (int pl = 0; pl < 2; ++pl)
        for (int ph = 0; ph < 2; ++ph) {
          int idx = q0_is_lo ? 2 * pl + ph : 2 * ph + pl;
          in[idx] = D[(a + cl * pl) + (size_t)M * (ph + 2 * c)];
        }
      for (int r = 0; r < 4; ++r) out[r] = m[4 * r] * in[0] + m[4 * r + 1] * in[1] + m[4 * r + 2] * in[2] + m[4 * r + 3] * in[3];
      for (int pl = 0; pl < 2; ++pl)
        for (int ph = 0; ph < 2; ++ph) {
          int idx = q0_is_lo ? 2 * pl + ph : 2 * ph + pl;
          T[(a + cl * pl) + (size_t)M * (ph + 2 * c)] = out[idx];
        }
    }
  // stabilizeTensorBody (:141-161)
  for (auto& x : T) if (std::abs(x) < 1e-100) x = 0.0;
  double t1 = now_s();
  o->t_gemm += t1 - t0;
  // step 3 SVD (:1552-1630): new bond = min(vol(lo)/ch, vol(hi)/ch)
  const int r = std::min(M, N);
  std::vector<double> S(r);
  std::vector<cplx> U((size_t)M * r), Vh((size_t)r * N);
  int info = 0, lwork = -1;
  cplx wq;
  std::vector<double> rwork;
  if (!p_zgesvd) return 2;
  if (o->use_gesdd) {
    size_t mn = r, mx = std::max(M, N);
    rwork.resize(std::max<size_t>(1, std::max(5 * mn * mn + 5 * mn, 2 * mx * mn + 2 * mn * mn + mn)));
    std::vector<int> iwork(8 * r);
    p_zgesdd("S", &M, &N, T.data(), &M, S.data(), U.data(), &M, Vh.data(), &r, &wq, &lwork, rwork.data(), iwork.data(), &info);
    lwork = (int)wq.real() + 1;
    std::vector<cplx> work(lwork);
    p_zgesdd("S", &M, &N, T.data(), &M, S.data(), U.data(), &M, Vh.data(), &r, work.data(), &lwork, rwork.data(), iwork.data(), &info);
  } else {
    rwork.resize(5 * r);
    p_zgesvd("S", "S", &M, &N, T.data(), &M, S.data(), U.data(), &M, Vh.data(), &r, &wq, &lwork, rwork.data(), &info);
    lwork = (int)wq.real() + 1;
    std::vector<cplx> work(lwork);
    p_zgesvd("S", "S", &M, &N, T.data(), &M, S.data(), U.data(), &M, Vh.data(), &r, work.data(), &lwork, rwork.data(), &info);
  }
  if (info != 0) return 3;
  double t2 = now_s();
  o->t_svd += t2 - t1;
  if (o->null_tol > 0) {
    double fro = 0;
    for (int k = 0; k < r; ++k) fro += S[k] * S[k];
    fro = std::sqrt(fro);
    for (int k = 1; k < r; ++k) if (S[k] <= o->null_tol * fro) S[k] = 0.0;
  }
  // truncateSvdTensors (:2366-2536).  Under the reference gauge both partial norms equal
  // sigma_k (sum of squares, comment :2421-2423) or sqrt(sigma_k) (true 2-norm).
  int cut = r;
  for (int k = 0; k < r; ++k) {
    double metric = o->cutoff_on_sqrt ? std::sqrt(S[k]) : S[k];
    if (metric < o->svd_cutoff) { cut = k + 1; break; }   // "return i + 1" at :2439
  }
  const int keep = std::max(1, std::min(cut, o->max_bond));   // :2445
  double tot = 0, kept = 0;
  for (int k = 0; k < r; ++k) { tot += S[k] * S[k]; if (k < keep) kept += S[k] * S[k]; }
  if (tot > 0) { const double w = (tot - kept) / tot; o->discarded += w; o->log_fidelity += std::log1p(-std::min(w, 1.0 - 1e-300)); }
  double rn = (o->renorm && kept > 0) ? std::sqrt(tot / kept) : 1.0;
  A.dr = keep; B.dl = keep;
  A.t.assign((size_t)M * keep, 0.0);
  B.t.assign((size_t)keep * N, 0.0);
  for (int k = 0; k < keep; ++k) {
    double fl, fr;
    if (o->gauge == 0) { fl = fr = std::sqrt(S[k]); }
    else if (o->gauge == 1) { fl = 1.0; fr = S[k]; }
    else { fl = S[k]; fr = 1.0; }
    fr *= rn;
    for (int i = 0; i < M; ++i) A.t[i + (size_t)M * k] = U[i + (size_t)M * k] * fl;
    for (int j = 0; j < N; ++j) B.t[k + (size_t)keep * j] = Vh[k + (size_t)r * j] * fr;
  }
  o->sv[lo].assign(S.begin(), S.begin() + keep);
  o->t_trunc += now_s() - t2;
  return 0;
}

extern "C" void oracle_bond_dims(const Oracle* o, int* out) { for (int i = 0; i + 1 < o->n; ++i) out[i] = o->s[i].dr; }
extern "C" int oracle_singular_values(const Oracle* o, int bond, double* out, int cap) {
  const auto& v = o->sv[bond];
  int c = std::min<int>(cap, v.size());
  for (int i = 0; i < c; ++i) out[i] = v[i];
  return (int)v.size();
}
extern "C" double oracle_discarded_weight(const Oracle* o) { return o->discarded; }
extern "C" double oracle_fidelity_estimate(const Oracle* o) { return std::exp(o->log_fidelity); }
extern "C" void oracle_times(const Oracle* o, double* out3) { out3[0] = o->t_gemm; out3[1] = o->t_svd; out3[2] = o->t_trunc; }
extern "C" void oracle_get_site(const Oracle* o, int k, double* out_ri, int* shape3) {
  const Site& s = o->s[k];
  shape3[0] = s.dl; shape3[1] = 2; shape3[2] = s.dr;
  if (out_ri) memcpy(out_ri, s.t.data(), s.t.size() * sizeof(cplx));
}
extern "C" void oracle_set_site(Oracle* o, int k, const double* in_ri, int dl, int dr) {
  Site& s = o->s[k];
  s.dl = dl; s.dr = dr;
  s.t.assign(reinterpret_cast<const cplx*>(in_ri), reinterpret_cast<const cplx*>(in_ri) + (size_t)2 * dl * dr);
}

// evaluateSync(ket) (:591-597): Root(i0..in-1), i0 fastest -> qubit 0 is the least-significant bit.
extern "C" int oracle_statevector(const Oracle* o, double* out_ri) {
  if (o->n > 28) return 1;
  std::vector<cplx> cur(o->s[0].t);   // (1*2) x dr0 viewed as (2 x dr0)
  size_t rows = 2;
  for (int k = 1; k < o->n; ++k) {
    const Site& s = o->s[k];
    std::vector<cplx> nxt(rows * 2 * s.dr);
    gemm((int)rows, 2 * s.dr, s.dl, cur.data(), (int)rows, s.t.data(), s.dl, nxt.data(), (int)rows);
    cur.swap(nxt);
    rows *= 2;
  }
  memcpy(out_ri, cur.data(), sizeof(cplx) * rows);
  return 0;
}

// Transfer-matrix sweep: <psi| prod_k O_k |psi>, O_k diagonal in Z basis with weights w[k][0], w[k][1].
// Same value as the reference's state-vector sums (:604-644) but usable for any n
// (algorithm of ITensorMPSVisitor.cpp:173-250 without the gauge move).
static cplx sweep_diag(const Oracle* o, const std::vector<std::array<double, 2>>& w) {
  std::vector<cplx> E{cplx(1, 0)};   // E[a', a], a' = bra index; 1x1
  for (int k = 0; k < o->n; ++k) {
    const Site& s = o->s[k];
    const int dl = s.dl, dr = s.dr;
    // F (2dl x dr) rows (a', p): F[(a',p), c] = w_p * sum_a E[a',a] A[a,p,c]
    std::vector<cplx> F((size_t)2 * dl * dr, 0.0), G((size_t)dl * dr);
    for (int p = 0; p < 2; ++p) {
      if (w[k][p] == 0.0) continue;
      // A_p as (dl x dr) with ld = 2*dl, offset p*dl
      gemm(dl, dr, dl, E.data(), dl, s.t.data() + (size_t)p * dl, 2 * dl, G.data(), dl);
      for (int c = 0; c < dr; ++c)
        for (int a = 0; a < dl; ++a) F[(a + dl * p) + (size_t)2 * dl * c] = w[k][p] * G[a + (size_t)dl * c];
    }
    // E'[c', c] = sum_{a',p} conj(A[a',p,c']) F[(a',p), c]
    std::vector<cplx> En((size_t)dr * dr);
    gemm(dr, dr, 2 * dl, s.t.data(), 2 * dl, F.data(), 2 * dl, En.data(), dr, 'C');
    E.swap(En);
  }
  return E[0];
}
extern "C" double oracle_norm(const Oracle* o) {
  std::vector<std::array<double, 2>> w(o->n, {1.0, 1.0});
  return sweep_diag(o, w).real();
}
// "exp-val-z" (:616-644): sum_x (-1)^{parity of x on measured bits} |amp_x|^2, NOT divided by the norm.
extern "C" double oracle_expval_z(const Oracle* o, const int* qubits, int nq) {
  std::vector<std::array<double, 2>> w(o->n, {1.0, 1.0});
  for (int i = 0; i < nq; ++i) w[qubits[i]][1] = -w[qubits[i]][1];   // Z^2 = I when listed twice
  return sweep_diag(o, w).real();
}
// <bits|psi> with fixed bits (0/1 for every qubit); computeWaveFuncSlice (:2588-2675) with no open leg.
extern "C" void oracle_amplitude(const Oracle* o, const int8_t* bits, double* out_ri) {
  std::vector<cplx> v{cplx(1, 0)};
  for (int k = 0; k < o->n; ++k) {
    const Site& s = o->s[k];
    std::vector<cplx> nv(s.dr, 0.0);
    for (int c = 0; c < s.dr; ++c)
      for (int a = 0; a < s.dl; ++a) nv[c] += v[a] * s.t[a + (size_t)s.dl * (bits[k] + 2 * c)];
    v.swap(nv);
  }
  out_ri[0] = v[0].real(); out_ri[1] = v[0].imag();
}

extern "C" void oracle_measure(Oracle* o, int q) { o->measure.push_back(q); }
extern "C" void oracle_clear_measure(Oracle* o) { o->measure.clear(); }

// GenerateSamples, utils/GateMatrixAlgebra.hpp:125-156 (n < 20 branch of finalize, :645-648).
// out: shots strings of nq chars (no terminator); returns the number of strings produced
// (can be < shots when the norm is < 1, cheat-sheet item 7 of SURVEY.md section 8a).
extern "C" int oracle_sample_statevector(Oracle* o, int shots, char* out) {
  const int nq = (int)o->measure.size();
  std::vector<cplx> sv((size_t)1 << o->n);
  oracle_statevector(o, reinterpret_cast<double*>(sv.data()));
  std::vector<double> rs;
  rs.reserve(shots + 1);
  for (int i = 0; i < shots; ++i) rs.push_back(std::uniform_real_distribution<double>(0.0, 1.0)(o->rng));   // RandomEngine.hpp:28-37
  std::sort(rs.begin(), rs.end());
  rs.push_back(2.0);   // the reference reads rs[m] one past the end inside reserved capacity; make that read defined
  double csum = 0.0;
  uint64_t m = 0;
  for (uint64_t k = 0; k < sv.size(); ++k) {
    csum += std::norm(sv[k]);
    while (m < (uint64_t)shots && rs[m] < csum) {
      for (int i = 0; i < nq; ++i) out[m * nq + i] = (k & (1ULL << o->measure[i])) ? '1' : '0';
      ++m;
    }
  }
  return (int)m;
}

// getMeasureSample, ExaTnMpsVisitor.cpp:2211-2364 (n >= 20 branch, :651-670): one shot.
// For each measured qubit (Measure order) the diagonal of the 1-qubit RDM of
// <psi| prod_prev (Pi_prev / p_prev) |psi> with that leg open; one uniform draw; 0 iff r <= p0.
extern "C" int oracle_sample_rdm_shot(Oracle* o, char* out) {
  const int nq = (int)o->measure.size();
  std::vector<int> res;
  std::vector<double> probs;
  for (int mi = 0; mi < nq; ++mi) {
    const int q = o->measure[mi];
    double pb[2];
    for (int b = 0; b < 2; ++b) {
      std::vector<std::array<double, 2>> w(o->n, {1.0, 1.0});
      for (size_t j = 0; j < res.size(); ++j) {
        // collapse tensor diag(1/p, 0) or diag(0, 1/p) appended to the ket only (:2236-2277)
        int qq = o->measure[j];
        w[qq][res[j]] *= 1.0 / probs[j];
        w[qq][1 - res[j]] = 0.0;
      }
      w[q][1 - b] = 0.0;   // open leg, diagonal element b (an already-collapsed qubit keeps its 1/p weight)
      pb[b] = sweep_diag(o, w).real();
    }
    const double PROB_EPS = 1e-12;   // :2329
    double p0 = std::fabs(pb[0]) < PROB_EPS ? 0.0 : pb[0];
    double p1 = std::fabs(pb[1]) < PROB_EPS ? 0.0 : pb[1];
    double r = std::uniform_real_distribution<double>(0.0, 1.0)(o->rng);
    int bit = (r <= p0) ? 0 : 1;
    res.push_back(bit);
    probs.push_back(bit == 0 ? p0 : p1);
    out[mi] = bit ? '1' : '0';
  }
  return 0;
}
