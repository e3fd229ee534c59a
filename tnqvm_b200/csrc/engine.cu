// Host engine of the B200 MPS gate-application path + the C ABI of include/mps_b200.h.
//
// What the reference does per gate through ~12 ExaTN tensor create/destroy calls, two string-parsed
// network rebuilds, three host copies of theta and >= 10 global syncs (ExaTnMpsVisitor.cpp:1387-1731),
// this engine does as: queue the gate -> group queued gates into dependency layers (gates of a layer
// touch disjoint sites) -> per layer four batched launches families on one stream:
//   theta GEMM+gate (DMMA)  ->  block-Jacobi sweeps  ->  sort/truncate  ->  gather + back-multiply GEMM.
// The only host synchronisation per layer is the read-back of the kept bond dimensions.
#include <algorithm>
#include <array>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <cfloat>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <cstdlib>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges show up in Nsight Systems / Compute timelines, no-ops otherwise

#include "../../include/mps_b200.h"
#include "kernels.h"

using namespace mpsb200;
typedef std::complex<double> cplx;

#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e__ = (call);                                                                                 \
    if (e__ != cudaSuccess)                                                                                   \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__));                                                     \
  } while (0)

namespace {

// NVTX range over a scope: the phases of the two-qubit step carry the reference's stat bucket names
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct SiteBuf {
  double2* d = nullptr;
  size_t cap = 0;   // complex elements
  int dl = 1, dr = 1;
};

struct QGate {
  int q0, q1;       // q1 < 0: single-qubit gate
  cplx m[16];
  int skip = 0;     // fuse_2q: the product of the merged gates is the identity, nothing to execute
};

// bump allocator over one grow-only device buffer
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void reset() { off = 0; }
  size_t reserve(size_t bytes) {   // returns offset
    size_t o = (off + 255) & ~size_t(255);
    off = o + bytes;
    return o;
  }
};

static std::string g_create_error;

// developer aid: MPS_B200_BACKTRACE=1 prints a native backtrace on SIGSEGV
static void segv_handler(int sig) {
  void* frames[64];
  int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}

// ---- sites sharded over the devices of one box (replaces the MPI site blocks of ExaTnMpsVisitor.cpp:347-531 and
// the boundary dispatch of :2059-2170).  One host process, one engine (sub-handle) and one worker thread per device.
// A boundary site travels as ONE peer copy over NVLink each way (32 chi^2 bytes); nothing else moves.
struct XferSlot {
  std::mutex mu;
  std::condition_variable cv;
  bool ready = false;
  const double2* ptr = nullptr;
  int dl = 0, dr = 0, src_dev = 0;
  cudaEvent_t ev = nullptr;   // recorded on the sender's stream behind the work that produced the tensor
};
struct ShardOp {
  enum Kind { LAYER, SEND, RECV } kind;
  int site = -1, slot = -1;
  std::vector<int> gates;   // LAYER: indices into the coordinator's queue
};
struct ShardGroup {
  std::vector<mps_b200_handle*> sub;           // one engine per device; sub[d] owns the sites [bounds[d].first, bounds[d].second)
  std::vector<std::pair<int, int>> bounds;
  std::vector<int> owner;                      // per site
  std::deque<XferSlot> slots;
  std::atomic<bool> abort{false};
  std::mutex err_mu;
  std::string first_error;
  double exchanges = 0, bytes_moved = 0;
  int gathered_ver = -1;                       // state version of the copy of all sites held by sub[0] (observables, v1)
};

}  // namespace

struct mps_b200_handle {
  ShardGroup* grp = nullptr;   // non-null: this handle is the coordinator of a site-sharded group and owns no device state
  std::vector<cudaEvent_t> xev;   // sub-handle of a group: events of its outgoing boundary transfers (reused across flushes)
  size_t xev_used = 0;
  std::vector<cudaEvent_t> oev;   // events ordering environment hand-overs between devices (observables of a group); round-robin
  size_t oev_next = 0;
  cudaEvent_t next_obs_event() {
    if (oev.size() < 32) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      oev.push_back(e);
      return e;
    }
    return oev[oev_next++ % oev.size()];
  }
  int nq = 0, nreg = 1, ntot = 0;
  int max_bond = INT_MAX - 1;
  double cutoff = DBL_MIN;
  int gauge = 0, device = 0;
  int cutoff_on_sqrt = 0, fuse_1q = 1, renorm = 0, profile = 0, layer_batch = 1, use_qr = 1;
  // post-SVD sanity check of ExaTnMpsVisitor.cpp:1632-1661 (both factors must have 2-norm >= 1e-3).  The reference prints
  // "[ERROR] Tensor norm validation failed!" and asserts, i.e. it crashes in Debug builds and carries on in Release builds (its
  // default, CMakeLists.txt:52-56).  2 (default) = Release behaviour: one line on stderr per handle, counted in mps_stats[12];
  // 1 = Debug behaviour: the call fails; 0 = off.
  int norm_guard = 2;
  // fuse_2q: consecutive 2q gates on the same site pair (and the 1q gates between them) become one 4x4 before they reach
  // the GPU, e.g. CX.Rz.CX = ZZ(gamma) of the QAOA circuits or Swap.Swap = 1 between two routed gates.
  // Off by default: the reference truncates after every 2q gate (ExaTnMpsVisitor.cpp:1394-1630), so with truncation
  // active the fused run is more accurate than, not identical to, the reference; without truncation both are exact.
  int fuse_2q = 0;
  std::vector<int> last_touch;   // per site: index in `queue` of the last queued gate on it (-1: none since the flush)
  double nfused2q = 0;
  double jacobi_tol = 0.0;   // 0 -> sqrt(M) * eps
  // Numerically-null components: sigma_k <= null_tol * ||theta||_F (<= 0: the default 1e-13) is rounding noise
  // of an exact zero.  Such columns are not rotated by the Jacobi sweeps and both factors of the component are zeroed at
  // write-back.  This is a DOCUMENTED DEVIATION from the reference, where LAPACK inside ExaTN returns noise singular values
  // (~1e-17 sigma_max) with arbitrary orthonormal vectors for rank-deficient thetas and the cut rule of :2434-2443 only drops
  // exact zeros; in the sqrt(S) gauge that noise grows by a square root per SVD until it competes with real weight for the
  // max-bond-dim slots (DESIGN.md section 1 quantifies it).  The engine cannot reproduce LAPACK's noise vectors (the second
  // factor comes from theta * G / sigma^2, which amplifies noise columns), so there is no "keep" mode; the CPU checker of
  // the test-suite mirrors this rule through an option of its own so that parity can be checked like for like.
  double null_tol = 0.0;
  int max_sweeps = 40;
  cudaStream_t stream = nullptr;
  std::vector<SiteBuf> sites;
  // VQE mode: the ansatz state every observable term starts from (mps_snapshot / mps_restore)
  std::vector<SiteBuf> snap;
  std::vector<std::vector<double>> snap_sv;
  double snap_discarded = 0.0, snap_log_fidelity = 0.0;
  bool has_snap = false;
  std::vector<char> has1q;
  std::vector<std::array<cplx, 4>> p1q;
  std::vector<QGate> queue;
  std::vector<std::vector<double>> sv;   // per bond
  std::vector<int> measure;
  std::mt19937_64 rng;
  double discarded = 0.0;
  double log_fidelity = 0.0;   // sum over truncations of log(1 - discarded/total): the fidelity estimate prod(1 - w) of config 5
  std::string err;
  // <psi|psi> of a register is remembered until its state changes (expval_z_all computes it on the way)
  uint64_t state_ver = 1;
  std::vector<uint64_t> norm_ver;
  std::vector<double> norm_val;
  // device workspace
  Arena ws;
  // pinned staging: two regions (pre-sync / post-sync uploads), plus read-back area
  char* pin[2] = {nullptr, nullptr};
  size_t pin_cap[2] = {0, 0};
  char* pin_rb = nullptr;
  size_t pin_rb_cap = 0;
  cudaEvent_t ev[6] = {};
  cudaEvent_t sev[2] = {};   // Jacobi sweep read-backs (the host runs one sweep behind the device)
  cudaStream_t stream2 = nullptr;   // right-to-left environment sweep of the observables, concurrent with the left-to-right one
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  int sm_count = 148;
  // persistent sweep kernel: resident CTAs per SM.  0 = by load: the routed circuits (configs 3 and 5) run layers of 2-8
  // gates, i.e. fewer pair tasks per tournament step than SMs x 4; launching only as many CTAs as there are tasks keeps
  // such tasks one per SM instead of letting up to four of them share an SM's FP64 pipe while other SMs idle
  int ctas_per_sm = 0;
  int wide_tasks = 1;      // 8 warps per pair task when at most two tasks per SM are resident (0: always 4)
  // Jacobi work matrices of one layer are processed in chunks whose total size fits this many MiB: a chunk stays
  // L2-resident over its sweeps (126 MB L2) instead of streaming from HBM at every tournament step.  chi=256 layers
  // (25 x 4 MiB) are one chunk, a chi=1024 theta (64 MiB) is a chunk of its own.
  int chunk_mb = 96;
  // 0 (default): the streaming sweep kernel (jacobi_svd.cu).  1: pair tasks resident in shared memory, split by rows over
  // thread-block clusters (jacobi_cluster.cu) -- parity-tested, but measured 2.5x SLOWER on the saturated chi=256 step
  // (144 vs 58 ms, profiles/r04h_*): a resident 16-column pair costs 128 KB of shared memory, so only 148 tasks are in flight
  // against 592 for the streaming kernel, and the per-task latency chain (dependency wait, bulk load, Gram, two cluster
  // barriers around the leader's rotation phase, update) is ~24 us.  Kept as an experiment, see DESIGN.md section 4.
  int jacobi_cluster = 0;
  std::map<std::pair<int, int>, int> cluster_cap;   // (cluster size, rows per CTA) -> co-resident clusters on this device
  int l2_persist = 1;      // access-policy window (persisting L2 lines) over the Jacobi matrices of the running chunk
  size_t l2_persist_max = 0, l2_window_max = 0;
  double guard_violations = 0;   // norm-guard failures seen with norm_guard = 2
  double nonconverged = 0;   // matrices that were still rotating after max_sweeps sweeps (truncated anyway; reported in mps_stats)
  // counters
  double n2q = 0, n1q = 0, nlayers = 0, nsweeps = 0, nlaunch = 0, ms_theta = 0, ms_svd = 0, ms_wb = 0, ms_qr = 0;

  // ------------------------------------------------------------------ memory helpers
  void ensure_ws(size_t bytes) {
    if (bytes <= ws.cap) return;
    CK(cudaStreamSynchronize(stream));
    if (ws.base) CK(cudaFree(ws.base));
    size_t ncap = std::max(bytes + bytes / 4, size_t(1) << 22);
    CK(cudaMalloc(&ws.base, ncap));
    ws.cap = ncap;
  }
  char* pinned(int which, size_t bytes) {
    if (bytes > pin_cap[which]) {
      CK(cudaStreamSynchronize(stream));
      if (pin[which]) CK(cudaFreeHost(pin[which]));
      size_t ncap = std::max(bytes * 2, size_t(1) << 16);
      CK(cudaMallocHost(&pin[which], ncap));
      pin_cap[which] = ncap;
    }
    return pin[which];
  }
  char* pinned_rb(size_t bytes) {
    if (bytes > pin_rb_cap) {
      CK(cudaStreamSynchronize(stream));
      if (pin_rb) CK(cudaFreeHost(pin_rb));
      size_t ncap = std::max(bytes * 2, size_t(1) << 16);
      CK(cudaMallocHost(&pin_rb, ncap));
      pin_rb_cap = ncap;
    }
    return pin_rb;
  }
  void ensure_site(int k, int dl, int dr, bool keep_data) {
    SiteBuf& s = sites[k];
    size_t need = (size_t)2 * dl * dr;
    if (need > s.cap) {
      size_t ncap = need;
      // grow geometrically but never beyond what max_bond allows
      if (max_bond < (1 << 20)) ncap = std::max(need, std::min((size_t)2 * max_bond * max_bond, need * 2));
      // stream-ordered allocation from the device's default pool (kept warm, see mps_create): growing a site costs
      // microseconds and no synchronisation, which matters while the bonds of a fresh |0...0> state double gate by gate
      double2* nd = nullptr;
      CK(cudaMallocAsync((void**)&nd, ncap * sizeof(double2), stream));
      if (s.d) {
        if (keep_data)
          CK(cudaMemcpyAsync(nd, s.d, (size_t)2 * s.dl * s.dr * sizeof(double2), cudaMemcpyDeviceToDevice, stream));
        CK(cudaFreeAsync(s.d, stream));
      }
      s.d = nd;
      s.cap = ncap;
    }
    s.dl = dl;
    s.dr = dr;
  }

  // ------------------------------------------------------------------ state
  void reset_state() {
    CK(cudaStreamSynchronize(stream));
    ++state_ver;
    queue.clear();
    std::fill(last_touch.begin(), last_touch.end(), -1);
    std::fill(has1q.begin(), has1q.end(), 0);
    measure.clear();
    discarded = 0.0;
    log_fidelity = 0.0;
    const cplx zero_state[2] = {cplx(1, 0), cplx(0, 0)};
    for (int k = 0; k < ntot; ++k) {
      ensure_site(k, 1, 1, false);
      CK(cudaMemcpyAsync(sites[k].d, zero_state, sizeof(zero_state), cudaMemcpyHostToDevice, stream));
    }
    CK(cudaStreamSynchronize(stream));
    for (auto& v : sv) v.assign(1, 1.0);
  }

  // remember / go back to the current state of all registers (device-to-device, stream-ordered; no host copy)
  void snapshot_state() {
    flush();
    snap.resize(ntot);
    for (int k = 0; k < ntot; ++k) {
      const SiteBuf& s = sites[k];
      SiteBuf& c = snap[k];
      const size_t need = (size_t)2 * s.dl * s.dr;
      if (need > c.cap) {
        if (c.d) CK(cudaFreeAsync(c.d, stream));
        CK(cudaMallocAsync((void**)&c.d, need * sizeof(double2), stream));
        c.cap = need;
      }
      c.dl = s.dl; c.dr = s.dr;
      CK(cudaMemcpyAsync(c.d, s.d, need * sizeof(double2), cudaMemcpyDeviceToDevice, stream));
    }
    snap_sv = sv;
    snap_discarded = discarded;
    snap_log_fidelity = log_fidelity;
    has_snap = true;
  }
  void restore_state() {
    if (!has_snap) throw std::runtime_error("mps_restore without mps_snapshot");
    flush();
    ++state_ver;
    for (int k = 0; k < ntot; ++k) {
      const SiteBuf& c = snap[k];
      ensure_site(k, c.dl, c.dr, false);
      CK(cudaMemcpyAsync(sites[k].d, c.d, (size_t)2 * c.dl * c.dr * sizeof(double2), cudaMemcpyDeviceToDevice, stream));
    }
    sv = snap_sv;
    discarded = snap_discarded;
    log_fidelity = snap_log_fidelity;
  }

  int reg_of(int q) const { return q / nq; }

  // ------------------------------------------------------------------ gate queue
  void push_1q(int q, const cplx* m) {
    if (q < 0 || q >= ntot) throw std::runtime_error("qubit index out of range");
    if (fuse_1q) {
      if (!has1q[q]) {
        has1q[q] = 1;
        p1q[q] = {m[0], m[1], m[2], m[3]};
      } else {   // new = m * old
        auto o = p1q[q];
        p1q[q] = {m[0] * o[0] + m[1] * o[2], m[0] * o[1] + m[1] * o[3], m[2] * o[0] + m[3] * o[2], m[2] * o[1] + m[3] * o[3]};
      }
    } else {
      QGate g;
      g.q0 = q; g.q1 = -1;
      for (int i = 0; i < 4; ++i) g.m[i] = m[i];
      last_touch[q] = (int)queue.size();
      queue.push_back(g);
      if (!layer_batch) flush();
    }
  }
  void push_2q(int q0, int q1, const cplx* m) {
    if (q0 < 0 || q1 < 0 || q0 >= ntot || q1 >= ntot) throw std::runtime_error("qubit index out of range");
    if (std::abs(q0 - q1) != 1) throw std::runtime_error("two-qubit gate on non-adjacent qubits (run the nearest-neighbour pass first)");
    if (reg_of(q0) != reg_of(q1)) throw std::runtime_error("two-qubit gate across registers");
    QGate g;
    g.q0 = q0; g.q1 = q1;
    for (int i = 0; i < 16; ++i) g.m[i] = m[i];
    if (fuse_1q && (has1q[q0] || has1q[q1])) {
      // m' = m * (P_q0 (x) P_q1) in the (q0,q1) index order
      cplx id[4] = {1, 0, 0, 1};
      const cplx* a = has1q[q0] ? p1q[q0].data() : id;
      const cplx* b = has1q[q1] ? p1q[q1].data() : id;
      cplx kp[16];
      for (int r0 = 0; r0 < 2; ++r0)
        for (int r1 = 0; r1 < 2; ++r1)
          for (int c0 = 0; c0 < 2; ++c0)
            for (int c1 = 0; c1 < 2; ++c1) kp[(2 * r0 + r1) * 4 + 2 * c0 + c1] = a[2 * r0 + c0] * b[2 * r1 + c1];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          cplx s = 0;
          for (int k = 0; k < 4; ++k) s += m[4 * r + k] * kp[4 * k + c];
          g.m[4 * r + c] = s;
        }
      has1q[q0] = has1q[q1] = 0;
    }
    if (fuse_2q && last_touch[q0] >= 0 && last_touch[q0] == last_touch[q1] && queue[last_touch[q0]].q1 >= 0) {
      // the last queued gate on both sites is a 2q gate on this very pair: prev <- g * prev (in prev's qubit order)
      QGate& pv = queue[last_touch[q0]];
      const bool same = (pv.q0 == q0);
      cplx t[16];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          cplx s = 0;
          for (int k = 0; k < 4; ++k) {
            const int rr = same ? r : ((r & 1) << 1 | (r >> 1)), kk = same ? k : ((k & 1) << 1 | (k >> 1));
            s += g.m[4 * rr + kk] * pv.m[4 * k + c];
          }
          t[4 * r + c] = s;
        }
      double off = 0;
      for (int i = 0; i < 16; ++i) { pv.m[i] = t[i]; off = std::max(off, std::abs(t[i] - cplx((i % 5) == 0 ? 1.0 : 0.0, 0.0))); }
      pv.skip = off < 1e-15;
      nfused2q += 1;
      return;
    }
    last_touch[q0] = last_touch[q1] = (int)queue.size();
    queue.push_back(g);
    if (!layer_batch) flush();
  }

  void group_flush();   // defined below the struct
  void flush() {
    if (grp) { group_flush(); return; }
    if (!queue.empty() || std::find(has1q.begin(), has1q.end(), (char)1) != has1q.end()) ++state_ver;
    if (!queue.empty()) {
      // dependency layering: a gate goes one layer after the last gate touching any of its sites
      std::vector<int> level(ntot, -1);
      std::vector<std::vector<int>> layers;
      for (size_t i = 0; i < queue.size(); ++i) {
        const QGate& g = queue[i];
        if (g.skip) continue;
        int l = level[g.q0];
        if (g.q1 >= 0) l = std::max(l, level[g.q1]);
        ++l;
        if ((int)layers.size() <= l) layers.resize(l + 1);
        layers[l].push_back((int)i);
        level[g.q0] = l;
        if (g.q1 >= 0) level[g.q1] = l;
      }
      for (auto& L : layers) run_layer(L);
      queue.clear();
      std::fill(last_touch.begin(), last_touch.end(), -1);
    }
    // leftover fused single-qubit gates
    std::vector<Gate1qProblem> p1;
    for (int q = 0; q < ntot; ++q)
      if (has1q[q]) {
        Gate1qProblem p;
        p.site = sites[q].d; p.dl = sites[q].dl; p.dr = sites[q].dr;
        for (int i = 0; i < 4; ++i) p.m[i] = make_double2(p1q[q][i].real(), p1q[q][i].imag());
        p1.push_back(p);
        has1q[q] = 0;
      }
    run_1q(p1);
  }

  void run_1q(const std::vector<Gate1qProblem>& p1) {
    if (p1.empty()) return;
    size_t bytes = p1.size() * sizeof(Gate1qProblem);
    ensure_ws(bytes + 256);
    ws.reset();
    size_t o = ws.reserve(bytes);
    // staging region 1, not 0: run_layer fills region 0 with its descriptors right after this call, while this upload may
    // still be in flight; region 1 is only ever written after a stream synchronisation (here and after the sweeps)
    char* st = pinned(1, bytes);
    CK(cudaStreamSynchronize(stream));   // staging reuse safety (1q-only flushes are rare)
    memcpy(st, p1.data(), bytes);
    CK(cudaMemcpyAsync(ws.base + o, st, bytes, cudaMemcpyHostToDevice, stream));
    long mx = 0;
    for (auto& p : p1) mx = std::max(mx, (long)p.dl * p.dr);
    launch_gate1q((const Gate1qProblem*)(ws.base + o), (int)p1.size(), mx, stream);
    nlaunch += 1;
    n1q += p1.size();
    CK(cudaGetLastError());
  }

  // ------------------------------------------------------------------ one dependency layer
  void run_layer(const std::vector<int>& L) {
    std::vector<Gate1qProblem> p1;
    std::vector<int> g2;
    for (int i : L) {
      const QGate& g = queue[i];
      if (g.q1 < 0) {
        Gate1qProblem p;
        p.site = sites[g.q0].d; p.dl = sites[g.q0].dl; p.dr = sites[g.q0].dr;
        for (int k = 0; k < 4; ++k) p.m[k] = make_double2(g.m[k].real(), g.m[k].imag());
        p1.push_back(p);
      } else g2.push_back(i);
    }
    run_1q(p1);
    if (g2.empty()) return;
    const int B = (int)g2.size();
    nlayers += 1;
    n2q += B;

    struct Dim { int lo, cl, ch, cr, M, N, tall, Mg, Ng, Mj; size_t oT, oG, oY, oV, oTq, oWd, oVer, oSig2, oSigma, oPerm, oSP, oSO, oKeep, oW; };
    std::vector<Dim> D(B);
    ws.reset();
    // pass 1: sizes
    for (int b = 0; b < B; ++b) {
      const QGate& g = queue[g2[b]];
      Dim& d = D[b];
      d.lo = std::min(g.q0, g.q1);
      d.cl = sites[d.lo].dl; d.ch = sites[d.lo].dr; d.cr = sites[d.lo + 1].dr;
      d.M = 2 * d.cl; d.N = 2 * d.cr;
      d.tall = d.M >= d.N;
      d.Mg = d.tall ? d.M : d.N;
      d.Ng = d.tall ? d.N : d.M;
      d.Mj = use_qr ? d.Ng : d.Mg;   // rows of the Jacobi work matrix: R^H is Ng x Ng
    }
    // descriptor block first, then sigma block (contiguous for one read-back), then matrices
    const size_t oGemm = ws.reserve(sizeof(GemmProblem) * B);
    const size_t oJac = ws.reserve(sizeof(JacobiProblem) * B);
    const size_t oTr = ws.reserve(sizeof(TruncProblem) * B);
    const size_t oGat = ws.reserve(sizeof(GatherProblem) * B);
    const size_t oGemm2 = ws.reserve(sizeof(GemmProblem) * B);
    const size_t oQr = ws.reserve(sizeof(QrProblem) * B);
    int pstride = 2;   // progress flags per matrix of the persistent sweep kernels: one per 8-column block (x one per cluster rank)
    for (int b = 0; b < B; ++b) pstride = std::max(pstride, ((D[b].Ng + 7) / 8 + 1) & ~1);
    pstride *= jacobi_cluster_progress_ints_per_block();
    const size_t oProg = ws.reserve(sizeof(int) * ((size_t)B * pstride + (size_t)(max_sweeps + 4) * (B + 1)));   // progress flags, then one task counter per (chunk, sweep)
    const size_t oActive = ws.reserve(sizeof(int) * 2 * (B + 2));   // per chunk: [0] matrices still rotating, [1..] their indices (within the chunk)
    const size_t oFlags = ws.reserve(sizeof(int) * (4 * B + 4) + sizeof(double) * B + 16);   // dirty[B], done[B], per chunk (remaining, fault), pad, fro2[B]
    const size_t oKeepBlk = ws.reserve((sizeof(int) + 2 * sizeof(double)) * B + 64);
    size_t sig_total = 0;
    for (int b = 0; b < B; ++b) sig_total += D[b].Ng;
    const size_t oSigmaBlk = ws.reserve(sizeof(double) * sig_total);
    {
      size_t so = 0;
      for (int b = 0; b < B; ++b) {
        Dim& d = D[b];
        d.oSigma = oSigmaBlk + sizeof(double) * so;
        so += d.Ng;
        d.oKeep = oKeepBlk + sizeof(int) * b;
        d.oW = oKeepBlk + ((sizeof(int) * B + 15) & ~size_t(15)) + 2 * sizeof(double) * b;
      }
    }
    // the Jacobi work matrices of the layer are contiguous (one L2 access-policy window per chunk)
    const size_t oGfirst = ws.reserve(0);
    for (int b = 0; b < B; ++b) D[b].oG = ws.reserve(sizeof(double2) * (size_t)D[b].Mj * D[b].Ng);
    for (int b = 0; b < B; ++b) {
      Dim& d = D[b];
      d.oWd = ws.reserve(sizeof(double2) * 64 * (size_t)((d.Ng + 7) / 8 + 1));
      {
        const size_t nbe_ = (size_t)((((d.Ng + 7) / 8) + 1) & ~1);
        d.oVer = ws.reserve(sizeof(int) * nbe_ + 8 + sizeof(int2) * nbe_ * nbe_);   // versions, then the clean-pair memo
      }
      d.oSig2 = ws.reserve(sizeof(double) * d.Ng);
      d.oPerm = ws.reserve(sizeof(int) * d.Ng);
      d.oSP = ws.reserve(sizeof(double) * d.Ng);
      d.oSO = ws.reserve(sizeof(double) * d.Ng);
      d.oT = ws.reserve(sizeof(double2) * (size_t)d.Mg * d.Ng);
      if (use_qr) {
        d.oY = ws.reserve(sizeof(double2) * (size_t)d.Mg * d.Ng);
        d.oV = ws.reserve(sizeof(double2) * 2 * (size_t)d.Mg * QR_PB);
        d.oTq = ws.reserve(sizeof(double2) * 2 * QR_PB * QR_PB);
      }
    }
    const size_t total = ws.off;
    ensure_ws(total);
    char* wb = ws.base;

    // Jacobi chunks: consecutive matrices whose work matrices fit `chunk_mb` together
    struct Chunk { int c0, c1, pairs, steps; long tasks; size_t bytes; int maxM; };
    std::vector<Chunk> chunks;
    {
      const size_t budget = (size_t)std::max(1, chunk_mb) << 20;
      for (int c0 = 0; c0 < B;) {
        Chunk c{c0, c0, 1, 1, 0, 0, 1};
        while (c.c1 < B) {
          const Dim& d = D[c.c1];
          const size_t sz = sizeof(double2) * (size_t)d.Mj * d.Ng;
          if (c.c1 > c0 && c.bytes + sz > budget) break;
          c.bytes += sz;
          const int nb = (d.Ng + 7) / 8, nbe = (nb == 1) ? 1 : ((nb + 1) & ~1);
          const int np = nb == 1 ? 1 : nbe / 2;
          c.pairs = std::max(c.pairs, np);
          c.maxM = std::max(c.maxM, d.Mj);
          c.steps = std::max(c.steps, nb == 1 ? 1 : nbe - 1);
          c.tasks += np;
          ++c.c1;
        }
        chunks.push_back(c);
        c0 = c.c1;
      }
    }

    // pass 2: descriptors
    const size_t descBytes = oFlags + sizeof(int) * (4 * B + 4) + sizeof(double) * B + 16 - oGemm;
    char* st = pinned(0, descBytes);
    memset(st, 0, descBytes);
    GemmProblem* hG = (GemmProblem*)(st + (oGemm - oGemm));
    JacobiProblem* hJ = (JacobiProblem*)(st + (oJac - oGemm));
    TruncProblem* hT = (TruncProblem*)(st + (oTr - oGemm));
    QrProblem* hQ = (QrProblem*)(st + (oQr - oGemm));
    int max_tiles = 0, max_pairs = 1, max_steps = 1, maxMg = 1, maxNg = 1;
    for (int b = 0; b < B; ++b) {
      const QGate& g = queue[g2[b]];
      const Dim& d = D[b];
      GemmProblem& p = hG[b];
      p.A = sites[d.lo].d; p.B = sites[d.lo + 1].d;
      p.C = (double2*)(wb + d.oT); p.C2 = (double2*)(wb + (use_qr ? d.oY : d.oG));
      p.M = d.cl; p.N = d.cr; p.K = d.ch;
      p.lda = 2 * d.cl; p.ldb = d.ch; p.ldc = d.Mg;
      p.b_col_stride = 1; p.b_col_off = 0;
      p.mode = 1; p.conjT_out = d.tall ? 0 : 1; p.alpha = 1.0;
      const bool q0lo = (g.q0 == d.lo);
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          // device index is 2*p_lo + p_hi; the gate's own index is 2*bit(q0)+bit(q1)  (ExaTnMpsVisitor.cpp:1492-1499)
          const int rr = q0lo ? r : ((r & 1) << 1 | (r >> 1));
          const int cc = q0lo ? c : ((c & 1) << 1 | (c >> 1));
          p.gate[r * 4 + c] = make_double2(g.m[rr * 4 + cc].real(), g.m[rr * 4 + cc].imag());
        }
      max_tiles = std::max(max_tiles, gemm_tiles(p.M, p.N, 1));
      JacobiProblem& j = hJ[b];
      j.G = (double2*)(wb + d.oG); j.M = d.Mj; j.N = d.Ng; j.ldg = d.Mj;
      if (use_qr) {
        QrProblem& q = hQ[b];
        q.Y = (double2*)(wb + d.oY); q.V = (double2*)(wb + d.oV); q.T = (double2*)(wb + d.oTq); q.G = j.G;
        q.M = d.Mg; q.N = d.Ng; q.ldy = d.Mg;
      }
      j.nb = (d.Ng + 7) / 8;
      j.nbe = (j.nb == 1) ? 1 : ((j.nb + 1) & ~1);
      j.wd = (double2*)(wb + d.oWd);
      {
        const size_t nbe_ = (size_t)((j.nb + 1) & ~1);
        j.ver = (int*)(wb + d.oVer);
        j.rec = (int2*)(wb + d.oVer + ((sizeof(int) * nbe_ + 7) & ~size_t(7)));
        CK(cudaMemsetAsync(wb + d.oVer, 0, sizeof(int) * nbe_ + 8 + sizeof(int2) * nbe_ * nbe_, stream));
      }
      max_pairs = std::max(max_pairs, j.nb == 1 ? 1 : j.nbe / 2);
      max_steps = std::max(max_steps, j.nb == 1 ? 1 : j.nbe - 1);
      maxMg = std::max(maxMg, d.Mg);
      maxNg = std::max(maxNg, d.Ng);
      TruncProblem& t = hT[b];
      // with the QR pre-reduction the Jacobi columns carry the singular vectors of the OTHER side (G = R^H)
      t.G = j.G; t.M = d.Mj; t.N = d.Ng; t.ldg = d.Mj; t.tall = use_qr ? !d.tall : d.tall;
      t.sig2 = (double*)(wb + d.oSig2); t.sigma = (double*)(wb + d.oSigma); t.perm = (int*)(wb + d.oPerm);
      t.scaleP = (double*)(wb + d.oSP); t.scaleO = (double*)(wb + d.oSO);
      t.keep = (int*)(wb + d.oKeep); t.weights = (double*)(wb + d.oW);
    }
    for (size_t k = 0; k < chunks.size(); ++k) {   // per chunk: "all its matrices rotating" (indices within the chunk)
      int* ha = (int*)(st + (oActive - oGemm)) + (chunks[k].c0 + 2 * k);
      ha[0] = chunks[k].c1 - chunks[k].c0;
      for (int b = 0; b < ha[0]; ++b) ha[1 + b] = b;
    }
    CK(cudaMemcpyAsync(wb + oGemm, st, descBytes, cudaMemcpyHostToDevice, stream));   // flags zeroed too
    int* d_dirty = (int*)(wb + oFlags);
    int* d_done = d_dirty + B;
    int* d_rem = d_done + B;   // per chunk c: d_rem[2c] = matrices still rotating, d_rem[2c+1] = dataflow faults
    double* d_fro2 = (double*)(wb + oFlags + ((sizeof(int) * (4 * B + 4) + 7) & ~size_t(7)));

    NvtxRange nvtx_layer("mps_b200: Two-qubit Gate layer");
    if (profile) CK(cudaEventRecord(ev[0], stream));
    nvtxRangePushA("Contract Two-Qubit Gate Tensor (theta GEMM + gate)");
    launch_gemm((const GemmProblem*)(wb + oGemm), B, max_tiles, 0, stream);
    nvtxRangePop();
    nlaunch += 1;
    if (profile) CK(cudaEventRecord(ev[1], stream));

    // ---- QR pre-reduction: theta_o = Q R, the Jacobi runs on G = R^H
    nvtxRangePushA("Decompose Tensor SVD (QR pre-reduction + Jacobi sweeps)");
    if (use_qr) {
      launch_qr((const QrProblem*)(wb + oQr), B, maxMg, maxNg, stream);
      nlaunch += qr_launch_count(maxNg);
      CK(cudaGetLastError());
    }
    if (profile) CK(cudaEventRecord(ev[4], stream));

    // ---- Jacobi sweeps
    const double eps = 2.220446049250313e-16;
    // relative rotation tolerance: sqrt(rows) eps, but never below sqrt(2048) eps = 1e-14 so that, up to chi = 1024, it does not
    // depend on which matrices share a layer (one GPU vs a sharded chain, layer_batch on / off give the same rotations)
    const double tol = jacobi_tol > 0 ? jacobi_tol : std::max(std::sqrt((double)maxMg), std::sqrt(2048.0)) * eps;
    const double tol2_base = tol * tol;
    // numerically-null threshold relative to ||theta||_F.  A CONSTANT (1e-13 ~ 10 sqrt(2048) eps), not a function of the layer:
    // a threshold that followed the largest matrix of the layer made the null decisions of a small chain-end gate depend on
    // what else happened to be batched with it -- and in the sqrt(S) gauge a borderline component that is kept instead of
    // dropped is amplified at the following gates, so the same circuit gave different truncated states on one
    // GPU and sharded over eight (bench.py circuit_sharded parity 0.24 in the norm, profiles/r04p_*)
    const double ntol = null_tol > 0 ? null_tol : 1e-13;
    const double dead2_base = ntol * ntol;
    const size_t remBytes = (sizeof(int) * 2 * (size_t)(max_sweeps + 4) * chunks.size() + 255) & ~size_t(255);
    int* h_rem = (int*)pinned_rb(remBytes + sizeof(int) * (B + 4) + sizeof(double) * (2 * B + sig_total) + 256);
    launch_fro2((const JacobiProblem*)(wb + oJac), B, d_fro2, stream);
    nlaunch += 1;
    int sweep = 0;
    static const bool trace = getenv("MPS_B200_TRACE") != nullptr;   // developer aid: per-sweep progress on stderr
    {
      // Chunks: consecutive matrices whose work matrices fit `chunk_mb` together run their sweeps to convergence before the
      // next chunk starts, so that a chunk is L2-resident across its sweeps.  Convergence is decided on the device
      // (jacobi_check_kernel marks finished matrices, the sweep kernel skips them); the host only learns how many matrices
      // are still rotating -- one sweep late: sweep k+1 is already queued when the count of sweep k is read, and a sweep
      // over a finished chunk is an empty launch.  No stream synchronisation inside the loop.
      int* d_prog = (int*)(wb + oProg);
      int* d_cnt = d_prog + (size_t)B * pstride;
      const JacobiProblem* dJ = (const JacobiProblem*)(wb + oJac);
      int nchunk = 0;
      for (const Chunk& ck : chunks) {
        const int c0 = ck.c0, c1 = ck.c1, cpairs = ck.pairs, csteps = ck.steps;
        const long ctasks = ck.tasks;
        const size_t bytes = ck.bytes;
        const int CB = c1 - c0;
        int* d_active = (int*)(wb + oActive) + (c0 + 2 * nchunk);
        int* d_crem = d_rem + 2 * nchunk;
        int* d_ccnt = d_cnt + (size_t)nchunk * (max_sweeps + 4);
        const bool window = l2_persist && l2_window_max > 0 && true;
        if (window) {
          cudaStreamAttrValue av;
          memset(&av, 0, sizeof(av));
          av.accessPolicyWindow.base_ptr = wb + D[c0].oG;
          av.accessPolicyWindow.num_bytes = std::min(bytes, l2_window_max);
          av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)l2_persist_max / (double)std::max<size_t>(1, av.accessPolicyWindow.num_bytes));
          av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          CK(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        int ccs = 0, crpc = 0, cmaxc = 0;
        bool use_cluster = jacobi_cluster && jacobi_cluster_shape(ck.maxM, &ccs, &crpc);
        if (use_cluster) {
          auto key = std::make_pair(ccs, crpc);
          auto it = cluster_cap.find(key);
          if (it == cluster_cap.end()) it = cluster_cap.emplace(key, jacobi_cluster_max_clusters(ccs, crpc)).first;
          cmaxc = it->second;
          if (cmaxc < 1) use_cluster = false;
        }
        int rotating = CB, csweep = 0, queued = 0, seen = 0;
        volatile int* h_crem = (volatile int*)(h_rem + 2 * (size_t)(max_sweeps + 4) * nchunk);
        auto enqueue = [&]() {
          // Safety net: a matrix still "rotating" after 20 sweeps is not converging slowly (quadratic convergence ended sweeps
          // ago), it sits at a rounding-noise floor; from sweep 20 on the tolerance is relaxed by 2x every second sweep, at most
          // 16x (<= 1e-13 at 2048 rows).  Converged matrices are unaffected, they no longer run.  (The stuck SVDs seen on the
          // real Sycamore circuit at chi = 512 were null columns rotating against each other, fixed by the dead-column
          // threshold, see jacobi_svd.cu.)
          const double relax = std::pow(4.0, (double)std::min(4, std::max(0, (queued - 18) / 2)));
          const double tol2 = tol2_base * relax;
          const double dead2 = dead2_base * relax;   // null columns a little above the base threshold stop rotating against each other
          if (use_cluster) {
            if (launch_jacobi_cluster_sweep(dJ + c0, rotating, cpairs, csteps, queued * csteps, tol2, dead2, d_fro2 + c0, d_dirty + c0, d_done + c0,
                                            d_prog + (size_t)c0 * pstride, pstride, d_crem + 1, d_active, ccs, crpc, cmaxc, stream) < 0)
              throw std::runtime_error(std::string("cluster Jacobi launch failed: ") + cudaGetErrorString(cudaGetLastError()));
            launch_jacobi_check(CB, d_dirty + c0, d_done + c0, d_crem, d_active, stream);
            nlaunch += 2;
            CK(cudaMemcpyAsync((void*)(h_crem + 2 * queued), d_crem, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaEventRecord(sev[queued & 1], stream));
            ++queued;
            return;
          }
          int per_sm = ctas_per_sm, warps = 4;
          // live tasks per tournament step (finished matrices have no slots: d_active); phantom slots of the smaller
          // matrices of the chunk still cost one dequeue each, so a CTA should not walk through more than ~100 per sweep
          const long est = std::max<long>(1, ctasks * rotating / CB);
          const long slots = (long)csteps * rotating * cpairs;
          const long want = std::max(est, slots / 96);
          if (per_sm <= 0) per_sm = (int)std::max<long>(1, std::min<long>(4, (want + sm_count - 1) / sm_count));
          // at most two tasks per SM: give each task eight warps (the register file holds 2 x 256 threads of this kernel)
          if (wide_tasks && per_sm <= 2) warps = 8;
          // at most one task per SM (a lone gate per layer in routed circuits, the straggler tail of a layer): sixteen warps --
          // the update phase is then one or two rows per thread instead of four, and that phase is a per-thread latency chain
          // (load 16 values, 64 dependent plane rotations, store): 14.7 of the 22.8 us a tournament step of a lone
          // 1024-row matrix takes (profiles/r04z_*)
          if (wide_tasks && ctas_per_sm <= 0 && est <= sm_count) { warps = 16; per_sm = 1; }
          launch_jacobi_sweep(dJ + c0, rotating, cpairs, csteps, queued * csteps, tol2, dead2, d_fro2 + c0, d_dirty + c0, d_done + c0,
                              d_prog + (size_t)c0 * pstride, pstride, d_ccnt + queued, d_crem + 1, sm_count * per_sm, warps, d_active, stream);
          launch_jacobi_check(CB, d_dirty + c0, d_done + c0, d_crem, d_active, stream);
          nlaunch += 2;
          CK(cudaMemcpyAsync((void*)(h_crem + 2 * queued), d_crem, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
          CK(cudaEventRecord(sev[queued & 1], stream));
          ++queued;
        };
        enqueue();
        for (;;) {
          if (queued < max_sweeps) enqueue();   // one sweep ahead of what the host has seen
          CK(cudaEventSynchronize(sev[seen & 1]));
          const int rem = h_crem[2 * seen], fault = h_crem[2 * seen + 1];
          ++seen;
          if (fault != 0) throw std::runtime_error("internal: Jacobi dataflow wait timed out");
          if (trace) fprintf(stderr, "[mps_b200 trace] layer %d chunk %d (%d matrices) sweep %d: %d still rotating\n", (int)nlayers, nchunk, CB, seen - 1, rem);
          if (rem == 0) { csweep = seen; break; }
          rotating = std::min(CB, std::max(1, rem));
          if (seen >= queued) {   // max_sweeps reached with matrices still rotating
            csweep = seen;
            nonconverged += rem;
            if (trace) {
              fprintf(stderr, "[mps_b200 trace] layer %d chunk %d NOT converged after %d sweeps: %d matrices; (lo site, M x N) of the chunk:", (int)nlayers, nchunk, seen, rem);
              for (int b = c0; b < c1; ++b) fprintf(stderr, " (%d, %d x %d)", D[b].lo, D[b].Mj, D[b].Ng);
              fprintf(stderr, "\n");
            }
            if (const char* dump = getenv("MPS_B200_DUMP_NONCONV")) {   // developer aid: the first stuck matrix (G as it is now, and theta_o) to a file
              static int dumped = 0;
              std::vector<int> hd(CB);
              CK(cudaStreamSynchronize(stream));
              CK(cudaMemcpy(hd.data(), d_done + c0, sizeof(int) * CB, cudaMemcpyDeviceToHost));
              for (int b = c0; b < c1 && dumped < 2; ++b)
                if (!hd[b - c0]) {
                  const Dim& d = D[b];
                  std::vector<double> buf(2 * (size_t)d.Mj * d.Ng), buf2(2 * (size_t)d.Mg * d.Ng);
                  CK(cudaMemcpy(buf.data(), wb + d.oG, sizeof(double2) * (size_t)d.Mj * d.Ng, cudaMemcpyDeviceToHost));
                  CK(cudaMemcpy(buf2.data(), wb + d.oT, sizeof(double2) * (size_t)d.Mg * d.Ng, cudaMemcpyDeviceToHost));
                  double f2 = 0;
                  CK(cudaMemcpy(&f2, d_fro2 + b, sizeof(double), cudaMemcpyDeviceToHost));
                  std::string fn = std::string(dump) + "_" + std::to_string(dumped) + ".bin";
                  if (FILE* f = fopen(fn.c_str(), "wb")) {
                    const double hdr[8] = {(double)d.Mj, (double)d.Ng, (double)d.Mg, tol, ntol, f2, (double)d.lo, (double)nlayers};
                    fwrite(hdr, sizeof(double), 8, f);
                    fwrite(buf.data(), sizeof(double), buf.size(), f);
                    fwrite(buf2.data(), sizeof(double), buf2.size(), f);
                    fclose(f);
                  }
                  ++dumped;
                }
            }
            break;
          }
        }
        sweep = std::max(sweep, csweep);
        ++nchunk;
        (void)c1;
      }
      if (l2_persist && l2_window_max > 0) {
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.num_bytes = 0;   // window off for the write-back
        av.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        CK(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av));
      }
    }
    nsweeps += sweep;
    if (profile) CK(cudaEventRecord(ev[2], stream));

    // ---- sort / truncate, read back kept dims + singular values
    nvtxRangePop();
    nvtxRangePushA("Truncate SVD Tensor (sort, cut rule, write-back)");
    launch_trunc((const TruncProblem*)(wb + oTr), B, cutoff, cutoff_on_sqrt, max_bond, gauge, renorm, ntol, stream);
    nlaunch += 1;
    const size_t keepBlkBytes = ((sizeof(int) * B + 15) & ~size_t(15)) + 2 * sizeof(double) * B;
    char* rb = (char*)h_rem + remBytes;
    CK(cudaMemcpyAsync(rb, wb + oKeepBlk, keepBlkBytes, cudaMemcpyDeviceToHost, stream));
    char* rbSig = rb + ((keepBlkBytes + 15) & ~size_t(15));
    CK(cudaMemcpyAsync(rbSig, wb + oSigmaBlk, sizeof(double) * sig_total, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const int* h_keep = (const int*)rb;
    const double* h_w = (const double*)(rb + ((sizeof(int) * B + 15) & ~size_t(15)));
    const double* h_sig = (const double*)rbSig;

    // ---- write-back
    const size_t gaBytes = (sizeof(GatherProblem) * B + 255) & ~size_t(255);   // GemmProblem needs 16-byte alignment
    char* st2 = pinned(1, gaBytes + sizeof(GemmProblem) * B);
    GatherProblem* hGa = (GatherProblem*)st2;
    GemmProblem* hG2 = (GemmProblem*)(st2 + gaBytes);
    memset(st2, 0, gaBytes + sizeof(GemmProblem) * B);
    int max_rows = 1, max_keep = 1, max_tiles2 = 1;
    size_t so = 0;
    for (int b = 0; b < B; ++b) {
      const Dim& d = D[b];
      const int keep = h_keep[b];
      if (keep < 1 || keep > d.Ng) throw std::runtime_error("internal: bad kept bond dimension");
      if (norm_guard) {
        // post-SVD sanity check, ExaTnMpsVisitor.cpp:1632-1661: ||Q_lo||_2 and ||Q_hi||_2 of the un-truncated factors must
        // both be >= 1e-3.  With the factors U f_lo(S) and f_hi(S) V^H these are sqrt(sum f(sigma_k)^2) -- known from the
        // singular values alone, no pass over the sites.
        double n_lo = 0.0, n_hi = 0.0;
        for (int k = 0; k < d.Ng; ++k) {
          const double s = h_sig[so + k];
          if (!(s > 0.0)) continue;
          n_lo += (gauge == 0) ? s : (gauge == 1 ? 1.0 : s * s);
          n_hi += (gauge == 0) ? s : (gauge == 1 ? s * s : 1.0);
        }
        if (std::sqrt(n_lo) < 1e-3 || std::sqrt(n_hi) < 1e-3) {
          if (guard_violations == 0 && norm_guard == 2)
            fprintf(stderr, "[ERROR] Tensor norm validation failed! sites (%d,%d): ||Q_lo|| = %g, ||Q_hi|| = %g (ExaTnMpsVisitor.cpp:1632-1661; reported once per handle)\n",
                    d.lo, d.lo + 1, std::sqrt(n_lo), std::sqrt(n_hi));
          guard_violations += 1;
          if (norm_guard == 1)
            throw std::runtime_error("tensor norm validation failed after the SVD on sites (" + std::to_string(d.lo) + "," + std::to_string(d.lo + 1) +
                                     "): ||Q_lo|| = " + std::to_string(std::sqrt(n_lo)) + ", ||Q_hi|| = " + std::to_string(std::sqrt(n_hi)) +
                                     " (ExaTnMpsVisitor.cpp:1632-1661; option norm_guard = 1 is the Debug-build behaviour)");
        }
      }
      if (h_w[2 * b] > 0) {
        const double w = (h_w[2 * b] - h_w[2 * b + 1]) / h_w[2 * b];
        discarded += w;
        log_fidelity += std::log1p(-std::min(w, 1.0 - 1e-300));
      }
      sv[d.lo].assign(h_sig + so, h_sig + so + keep);
      so += d.Ng;
      ensure_site(d.lo, d.cl, keep, false);
      ensure_site(d.lo + 1, keep, d.cr, false);
      GatherProblem& ga = hGa[b];
      GemmProblem& p = hG2[b];
      ga.G = (const double2*)(wb + d.oG); ga.perm = (const int*)(wb + d.oPerm); ga.scale = (const double*)(wb + d.oSP);
      ga.M = d.Mj; ga.keep = keep; ga.ldg = d.Mj;
      p.alpha = 1.0; p.mode = 0; p.conjT_out = 0; p.b_col_stride = 1; p.b_col_off = 0;
      if (use_qr) {
        // G = X' = V_o S (Ng x Ng): right singular vectors of theta_o (= T, Mg x Ng) times sigma
        p.A = (const double2*)(wb + d.oT); p.lda = d.Mg;
        p.B = (const double2*)(wb + d.oG); p.ldb = d.Mj; p.b_gather = (const int*)(wb + d.oPerm);
        p.M = d.Mg; p.N = keep; p.K = d.Ng;
        p.col_scale = (const double*)(wb + d.oSO);
        if (d.tall) {
          // theta = T:  hi = f_hi(s)/s * X'[:,perm]^H (keep x N);  lo = T X'[:,perm] diag(f_lo(s)/s^2)  (M x keep)
          ga.out = sites[d.lo + 1].d; ga.ldo = keep; ga.conjT = 1;
          p.C = sites[d.lo].d; p.ldc = d.M;
        } else {
          // theta^H = T:  lo = X'[:,perm] f_lo(s)/s (M x keep);  hi = (T X'[:,perm] diag(f_hi(s)/s^2))^H  (keep x N)
          ga.out = sites[d.lo].d; ga.ldo = d.M; ga.conjT = 0;
          p.C = sites[d.lo + 1].d; p.ldc = keep; p.conjT_out = 1;
        }
      } else if (d.tall) {
        // lo = G[:,perm] * sP  (M x keep);  hi = diag(sO) * G[:,perm]^H * theta  (keep x N)
        ga.out = sites[d.lo].d; ga.ldo = d.M; ga.conjT = 0;
        p.A = (const double2*)(wb + d.oG); p.lda = d.Mg; p.a_gather = (const int*)(wb + d.oPerm);
        p.B = (const double2*)(wb + d.oT); p.ldb = d.Mg;
        p.C = sites[d.lo + 1].d; p.ldc = keep;
        p.M = keep; p.N = d.N; p.K = d.Mg;
        p.row_scale = (const double*)(wb + d.oSO);
      } else {
        // hi = (G[:,perm] * sP)^H  (keep x N);  lo = theta * G[:,perm] * diag(sO) = T^H G[:,perm] diag(sO)  (M x keep)
        ga.out = sites[d.lo + 1].d; ga.ldo = keep; ga.conjT = 1;
        p.A = (const double2*)(wb + d.oT); p.lda = d.Mg;
        p.B = (const double2*)(wb + d.oG); p.ldb = d.Mg; p.b_gather = (const int*)(wb + d.oPerm);
        p.C = sites[d.lo].d; p.ldc = d.M;
        p.M = d.M; p.N = keep; p.K = d.Mg;
        p.col_scale = (const double*)(wb + d.oSO);
      }
      max_rows = std::max(max_rows, d.Mj);
      max_keep = std::max(max_keep, keep);
      max_tiles2 = std::max(max_tiles2, gemm_tiles(p.M, p.N, 0));
    }
    CK(cudaMemcpyAsync(wb + oGat, st2, sizeof(GatherProblem) * B, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(wb + oGemm2, st2 + gaBytes, sizeof(GemmProblem) * B, cudaMemcpyHostToDevice, stream));
    launch_gather((const GatherProblem*)(wb + oGat), B, max_rows, max_keep, stream);
    launch_gemm((const GemmProblem*)(wb + oGemm2), B, max_tiles2, use_qr ? 0 : 1, stream);
    nlaunch += 2;
    nvtxRangePop();
    CK(cudaGetLastError());
    if (profile) {
      CK(cudaEventRecord(ev[3], stream));
      CK(cudaEventSynchronize(ev[3]));
      float a = 0, b2 = 0, c = 0, q = 0;
      CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
      CK(cudaEventElapsedTime(&b2, ev[1], ev[2]));
      CK(cudaEventElapsedTime(&c, ev[2], ev[3]));
      CK(cudaEventElapsedTime(&q, ev[1], ev[4]));
      ms_theta += a; ms_svd += b2; ms_wb += c; ms_qr += q;
    }
  }

  // ------------------------------------------------------------------ observables
  // <psi| prod_k diag(w[k][0], w[k][1]) |psi> over the sites [s0, s1) by a left-to-right transfer sweep.
  // If envs != nullptr the environment before each site is kept there (device pointers into the arena).
  struct EnvBufs { std::vector<double2*> e; };

  void left_step(const SiteBuf& S, const double2* E, double2* F, double2* Eout, double w0, double w1, cudaStream_t st_ = nullptr) {
    cudaStream_t stream = st_ ? st_ : this->stream;   // shadows the member: every launch below goes to the chosen stream
    const int dl = S.dl, dr = S.dr;
    if (w0 == w1) {
      // both physical slices in one GEMM: the site is a column-major dl x (2 dr) matrix with column index p + 2c
      GemmProblem g;
      memset(&g, 0, sizeof(g));
      g.A = E; g.lda = dl; g.B = S.d; g.ldb = dl; g.C = F; g.ldc = dl;
      g.M = dl; g.N = 2 * dr; g.K = dl; g.b_col_stride = 1; g.alpha = w0;
      launch_gemm1(g, 0, stream);
      nlaunch += 1;
    } else
    for (int p = 0; p < 2; ++p) {
      GemmProblem g;
      memset(&g, 0, sizeof(g));
      g.A = E; g.lda = dl; g.B = S.d + (size_t)p * dl; g.ldb = 2 * dl; g.C = F + (size_t)p * dl; g.ldc = 2 * dl;
      g.M = dl; g.N = dr; g.K = dl; g.b_col_stride = 1; g.alpha = p ? w1 : w0;
      launch_gemm1(g, 0, stream);
    }
    GemmProblem g;
    memset(&g, 0, sizeof(g));
    g.A = S.d; g.lda = 2 * dl; g.B = F; g.ldb = 2 * dl; g.C = Eout; g.ldc = dr;
    g.M = dr; g.N = dr; g.K = 2 * dl; g.b_col_stride = 1; g.alpha = 1.0;
    launch_gemm1(g, 1, stream);
    nlaunch += (w0 == w1) ? 1 : 3;
  }
  // R_k[a,a'] = sum_{p,c,c'} A[a,p,c] R[c,c'] conj(A[a',p,c'])
  void right_step(const SiteBuf& S, const double2* R, double2* H, double2* Rout, cudaStream_t st_ = nullptr, double w0 = 1.0, double w1 = 1.0) {
    cudaStream_t stream = st_ ? st_ : this->stream;
    const int dl = S.dl, dr = S.dr;
    if (w0 != 1.0 || w1 != 1.0) {
      // H_p = w_p S_p R, one GEMM per physical slice
      for (int p = 0; p < 2; ++p) {
        GemmProblem g;
        memset(&g, 0, sizeof(g));
        g.A = S.d + (size_t)p * dl; g.lda = 2 * dl; g.B = R; g.ldb = dr; g.C = H + (size_t)p * dl; g.ldc = 2 * dl;
        g.M = dl; g.N = dr; g.K = dr; g.b_col_stride = 1; g.alpha = p ? w1 : w0;
        launch_gemm1(g, 0, stream);
      }
      nlaunch += 1;
    } else {
      // H[(a,p), c'] = sum_c S[(a,p), c] R[c, c']: the site as a (2 dl) x dr matrix, both physical slices in one GEMM
      GemmProblem g;
      memset(&g, 0, sizeof(g));
      g.A = S.d; g.lda = 2 * dl; g.B = R; g.ldb = dr; g.C = H; g.ldc = 2 * dl;
      g.M = 2 * dl; g.N = dr; g.K = dr; g.b_col_stride = 1; g.alpha = 1.0;
      launch_gemm1(g, 0, stream);
    }
    GemmProblem g;
    memset(&g, 0, sizeof(g));
    g.A = H; g.lda = dl; g.B = S.d; g.ldb = dl; g.C = Rout; g.ldc = dl;
    g.M = dl; g.N = dl; g.K = 2 * dr; g.b_col_stride = 1; g.alpha = 1.0;
    launch_gemm1(g, 2, stream);
    nlaunch += 2;
  }

  size_t max_env_elems(int s0, int s1) const {
    size_t m = 1;
    for (int k = s0; k < s1; ++k) m = std::max(m, (size_t)sites[k].dr * sites[k].dr);
    return m;
  }
  size_t max_site_elems(int s0, int s1) const {
    size_t m = 2;
    for (int k = s0; k < s1; ++k) m = std::max(m, (size_t)2 * sites[k].dl * sites[k].dr);
    return m;
  }

  cplx sweep_weights(int reg, const std::vector<std::array<double, 2>>& w) {
    flush();
    const int s0 = reg * nq, s1 = s0 + nq;
    const size_t me = max_env_elems(s0, s1), ms = max_site_elems(s0, s1);
    ws.reset();
    const size_t oE0 = ws.reserve(me * 16), oE1 = ws.reserve(me * 16), oF = ws.reserve(ms * 16);
    ensure_ws(ws.off);
    double2* E[2] = {(double2*)(ws.base + oE0), (double2*)(ws.base + oE1)};
    double2* F = (double2*)(ws.base + oF);
    const cplx one(1, 0);
    CK(cudaMemcpyAsync(E[0], &one, 16, cudaMemcpyHostToDevice, stream));
    int cur = 0;
    for (int k = s0; k < s1; ++k) {
      left_step(sites[k], E[cur], F, E[cur ^ 1], w[k - s0][0], w[k - s0][1]);
      cur ^= 1;
    }
    cplx out;
    CK(cudaMemcpyAsync(&out, E[cur], 16, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    return out;
  }

  // left environments L[k] (before site s0+k, k = 0..n) and right environments R[k] (after site s0+k-1)
  void build_envs(int reg, std::vector<double2*>& Lv, std::vector<double2*>& Rv, double2*& F, double2*& Etmp, double2*& Etmp2, double2*& scal, int nscal = 0) {
    flush();
    const int s0 = reg * nq, n = nq;
    ws.reset();
    std::vector<size_t> oL(n + 1), oR(n + 1);
    for (int k = 0; k <= n; ++k) {
      const size_t d = (k == 0) ? 1 : sites[s0 + k - 1].dr;
      oL[k] = ws.reserve(d * d * 16);
      oR[k] = ws.reserve(d * d * 16);
    }
    const size_t oF = ws.reserve(max_site_elems(s0, s0 + n) * 16);
    const size_t oF2 = ws.reserve(max_site_elems(s0, s0 + n) * 16);
    const size_t oT = ws.reserve(max_env_elems(s0, s0 + n) * 16);
    const size_t oT2 = ws.reserve(max_env_elems(s0, s0 + n) * 16);
    const size_t oS = ws.reserve(16 * (size_t)(n + 8 + nscal));
    ensure_ws(ws.off);
    Lv.resize(n + 1); Rv.resize(n + 1);
    for (int k = 0; k <= n; ++k) { Lv[k] = (double2*)(ws.base + oL[k]); Rv[k] = (double2*)(ws.base + oR[k]); }
    F = (double2*)(ws.base + oF);
    double2* F2 = (double2*)(ws.base + oF2);
    Etmp = (double2*)(ws.base + oT);
    Etmp2 = (double2*)(ws.base + oT2);
    scal = (double2*)(ws.base + oS);
    const cplx one(1, 0);
    CK(cudaMemcpyAsync(Lv[0], &one, 16, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(Rv[n], &one, 16, cudaMemcpyHostToDevice, stream));
    // the two environment chains are independent: left-to-right on the handle's stream, right-to-left on the second one
    CK(cudaEventRecord(fork_ev, stream));
    CK(cudaStreamWaitEvent(stream2, fork_ev, 0));
    for (int k = 0; k < n; ++k) left_step(sites[s0 + k], Lv[k], F, Lv[k + 1], 1.0, 1.0);
    for (int k = n - 1; k >= 0; --k) right_step(sites[s0 + k], Rv[k + 1], F2, Rv[k], stream2);
    CK(cudaEventRecord(join_ev, stream2));
    CK(cudaStreamWaitEvent(stream, join_ev, 0));
  }

  // <Z_k> for every k and the norm from ONE left and ONE right transfer sweep: the left sweep keeps F_k = L_k S_k, the
  // right sweep keeps H_k = S_k R_{k+1}, and <psi| Z_k |psi> = sum_{a,p,c} (-1)^p conj(F_k[a,p,c]) H_k[a,p,c]
  // (L_k is Hermitian).  The two sweeps are independent chains of small GEMMs (latency-bound along the chain), so they run
  // concurrently on two streams; all n reductions are then ONE batched multi-CTA launch and one read-back.
  void expval_z_all(int reg, double* out) {
    flush();
    const int s0 = reg * nq, n = nq;
    ws.reset();
    std::vector<size_t> oF(n), oHk(n);
    size_t maxL = 1;
    for (int k = 0; k < n; ++k) {
      const size_t bytes = (size_t)2 * sites[s0 + k].dl * sites[s0 + k].dr * 16;
      oF[k] = ws.reserve(bytes);
      oHk[k] = ws.reserve(bytes);
      maxL = std::max(maxL, (size_t)sites[s0 + k].dr * sites[s0 + k].dr);
    }
    constexpr int DOT_CTAS = 32;
    const size_t oE0 = ws.reserve(maxL * 16), oE1 = ws.reserve(maxL * 16), oR0 = ws.reserve(maxL * 16), oR1 = ws.reserve(maxL * 16);
    const size_t oS = ws.reserve(16 * (size_t)(n + 8));
    const size_t oP = ws.reserve(sizeof(DotProblem) * (size_t)n);
    const size_t oPart = ws.reserve(16 * (size_t)n * DOT_CTAS);
    ensure_ws(ws.off);
    double2* E[2] = {(double2*)(ws.base + oE0), (double2*)(ws.base + oE1)};
    double2* R[2] = {(double2*)(ws.base + oR0), (double2*)(ws.base + oR1)};
    double2* scal = (double2*)(ws.base + oS);
    const cplx one(1, 0);
    CK(cudaMemcpyAsync(E[0], &one, 16, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(R[0], &one, 16, cudaMemcpyHostToDevice, stream));
    std::vector<DotProblem> hp(n);
    for (int k = 0; k < n; ++k) {
      const SiteBuf& S = sites[s0 + k];
      hp[k].F = (const double2*)(ws.base + oF[k]); hp[k].H = (const double2*)(ws.base + oHk[k]);
      hp[k].total = 2L * S.dl * S.dr; hp[k].n = S.dl; hp[k].mode = 0; hp[k].w0 = 1.0; hp[k].w1 = -1.0;
    }
    CK(cudaMemcpyAsync(ws.base + oP, hp.data(), sizeof(DotProblem) * (size_t)n, cudaMemcpyHostToDevice, stream));   // pageable source: staged before the call returns
    CK(cudaEventRecord(fork_ev, stream));
    CK(cudaStreamWaitEvent(stream2, fork_ev, 0));
    int cur = 0;
    for (int k = 0; k < n; ++k) {
      left_step(sites[s0 + k], E[cur], (double2*)(ws.base + oF[k]), E[cur ^ 1], 1.0, 1.0);
      cur ^= 1;
    }
    CK(cudaMemcpyAsync(scal + n, E[cur], 16, cudaMemcpyDeviceToDevice, stream));   // <psi|psi>
    int rc = 0;
    for (int k = n - 1; k >= 0; --k) {
      right_step(sites[s0 + k], R[rc], (double2*)(ws.base + oHk[k]), R[rc ^ 1], stream2);
      rc ^= 1;
    }
    CK(cudaEventRecord(join_ev, stream2));
    CK(cudaStreamWaitEvent(stream, join_ev, 0));
    launch_dot_batch((const DotProblem*)(ws.base + oP), n, DOT_CTAS, (double2*)(ws.base + oPart), scal, stream);
    nlaunch += 2;
    std::vector<cplx> h(n + 1);
    CK(cudaMemcpyAsync(h.data(), scal, 16 * (size_t)(n + 1), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    for (int k = 0; k < n; ++k) out[k] = h[k].real();
    norm_val[reg] = h[n].real();
    norm_ver[reg] = state_ver;
  }

  void expval_zz_pairs(int reg, int np, const int* qi, const int* qj, double* out) {
    std::vector<double2*> Lv, Rv;
    double2 *F, *Et, *Et2, *scal;
    for (int t = 0; t < np; ++t)
      if (std::min(qi[t], qj[t]) < 0 || std::max(qi[t], qj[t]) >= nq) throw std::runtime_error("qubit index out of range");
    build_envs(reg, Lv, Rv, F, Et, Et2, scal, np);
    const int s0 = reg * nq;
    for (int t = 0; t < np; ++t) {
      const int i = std::min(qi[t], qj[t]), j = std::max(qi[t], qj[t]);
      if (i == j) launch_trace_pair(Lv[nq], Rv[nq], 1, scal + t, stream);
      else {
        left_step(sites[s0 + i], Lv[i], F, Et, 1.0, -1.0);
        double2* cur = Et; double2* nxt = Et2;
        for (int k = i + 1; k < j; ++k) { left_step(sites[s0 + k], cur, F, nxt, 1.0, 1.0); std::swap(cur, nxt); }
        left_step(sites[s0 + j], cur, F, nxt, 1.0, -1.0);
        launch_trace_pair(nxt, Rv[j + 1], sites[s0 + j].dr, scal + t, stream);
      }
      nlaunch += 1;
    }
    std::vector<cplx> h(np);   // ONE read-back for all pairs
    CK(cudaMemcpyAsync(h.data(), scal, 16 * (size_t)np, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    for (int t = 0; t < np; ++t) out[t] = h[t].real();
  }

  // getMeasureSample (ExaTnMpsVisitor.cpp:2211-2364) for registers of 20 qubits and more: per shot, the measured qubits in
  // Measure order; for each one the diagonal of its reduced density matrix conditioned on the outcomes so far, one uniform
  // draw, 0 iff r <= p0 (:2329-2345).  The reference contracts the whole <psi|psi> ladder for every measured qubit of every shot.
  // Here the unconditioned left / right environments are built once; a shot only re-propagates the environments its own
  // projectors invalidate (measuring in ascending qubit order: ONE transfer step per qubit, and H_q = S_q R_{q+1} is shared by
  // all shots).  p_b = sum conj(F_b) H_b with F = L_q S_q, H = S_q R_{q+1}.  Same draws in the same order as the reference.
  int sample_rdm(int reg, int shots, char* out) {
    flush();
    const int s0 = reg * nq, n = nq, nm = (int)measure.size();
    ws.reset();
    std::vector<size_t> oL0(n + 1), oR0(n + 1), oLc(n + 1), oRc(n + 1), oHb(n);
    for (int k = 0; k <= n; ++k) {
      const size_t d = (k == 0) ? 1 : sites[s0 + k - 1].dr;
      oL0[k] = ws.reserve(d * d * 16); oR0[k] = ws.reserve(d * d * 16);
      oLc[k] = ws.reserve(d * d * 16); oRc[k] = ws.reserve(d * d * 16);
    }
    std::vector<char> is_meas(n, 0);
    for (int q : measure) is_meas[q] = 1;
    for (int k = 0; k < n; ++k) oHb[k] = is_meas[k] ? ws.reserve((size_t)2 * sites[s0 + k].dl * sites[s0 + k].dr * 16) : 0;
    const size_t site_max = max_site_elems(s0, s0 + n) * 16;
    const size_t oF = ws.reserve(site_max), oF2 = ws.reserve(site_max), oH = ws.reserve(site_max);
    constexpr int DOT_CTAS = 16;
    const size_t oP = ws.reserve(sizeof(DotProblem) * (size_t)(4 * n));
    const size_t oPart = ws.reserve(16 * 2 * DOT_CTAS);
    const size_t oS = ws.reserve(64);
    ensure_ws(ws.off);
    auto env = [&](const std::vector<size_t>& o, int k) { return (double2*)(ws.base + o[k]); };
    double2* F = (double2*)(ws.base + oF);
    double2* F2 = (double2*)(ws.base + oF2);
    double2* Hs = (double2*)(ws.base + oH);
    double2* scal = (double2*)(ws.base + oS);
    const cplx one(1, 0);
    CK(cudaMemcpyAsync(env(oL0, 0), &one, 16, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(env(oR0, n), &one, 16, cudaMemcpyHostToDevice, stream));
    // dot descriptors: per site (p0, p1) against the shared H_q and against the per-shot scratch H
    std::vector<DotProblem> hp(4 * (size_t)n);
    for (int k = 0; k < n; ++k)
      for (int v = 0; v < 2; ++v)
        for (int b = 0; b < 2; ++b) {
          DotProblem& d = hp[4 * k + 2 * v + b];
          d.F = F; d.H = v ? Hs : (double2*)(ws.base + oHb[k]);
          d.total = 2L * sites[s0 + k].dl * sites[s0 + k].dr; d.n = sites[s0 + k].dl; d.mode = 0;
          d.w0 = b ? 0.0 : 1.0; d.w1 = b ? 1.0 : 0.0;
        }
    CK(cudaMemcpyAsync(ws.base + oP, hp.data(), sizeof(DotProblem) * hp.size(), cudaMemcpyHostToDevice, stream));
    // unconditioned environments (two concurrent chains); the right chain also leaves H_q = S_q R_{q+1} for the measured sites
    CK(cudaEventRecord(fork_ev, stream));
    CK(cudaStreamWaitEvent(stream2, fork_ev, 0));
    for (int k = 0; k < n; ++k) left_step(sites[s0 + k], env(oL0, k), F, env(oL0, k + 1), 1.0, 1.0);
    for (int k = n - 1; k >= 0; --k)
      right_step(sites[s0 + k], env(oR0, k + 1), is_meas[k] ? (double2*)(ws.base + oHb[k]) : F2, env(oR0, k), stream2);
    CK(cudaEventRecord(join_ev, stream2));
    CK(cudaStreamWaitEvent(stream, join_ev, 0));

    std::vector<const double2*> curL(n + 1), curR(n + 1);
    std::vector<std::array<double, 2>> w(n);
    cplx* h2 = (cplx*)pinned_rb(64);   // pinned: the two probabilities come back without a staging copy
    const double PROB_EPS = 1e-12;
    for (int s = 0; s < shots; ++s) {
      for (int k = 0; k <= n; ++k) { curL[k] = env(oL0, k); curR[k] = env(oR0, k); }
      std::fill(w.begin(), w.end(), std::array<double, 2>{1.0, 1.0});
      int Lvalid = n, Rvalid = 0;   // curL[0..Lvalid] and curR[Rvalid..n] are consistent with the projectors set so far
      for (int mi = 0; mi < nm; ++mi) {
        const int q = measure[mi];
        const SiteBuf& S = sites[s0 + q];
        for (int k = Lvalid; k < q; ++k) {
          left_step(sites[s0 + k], curL[k], F, env(oLc, k + 1), w[k][0], w[k][1]);
          curL[k + 1] = env(oLc, k + 1);
        }
        Lvalid = std::max(Lvalid, q);
        for (int k = Rvalid - 1; k > q; --k) {
          right_step(sites[s0 + k], curR[k + 1], F2, env(oRc, k), nullptr, w[k][0], w[k][1]);
          curR[k] = env(oRc, k);
        }
        Rvalid = std::min(Rvalid, q + 1);
        const bool shared_h = (curR[q + 1] == env(oR0, q + 1)) && w[q][0] == 1.0 && w[q][1] == 1.0;
        {   // F = L_q S_q (both physical slices in one GEMM)
          GemmProblem g;
          memset(&g, 0, sizeof(g));
          g.A = curL[q]; g.lda = S.dl; g.B = S.d; g.ldb = S.dl; g.C = F; g.ldc = S.dl;
          g.M = S.dl; g.N = 2 * S.dr; g.K = S.dl; g.b_col_stride = 1; g.alpha = 1.0;
          launch_gemm1(g, 0, stream);
          nlaunch += 1;
        }
        if (!shared_h) {   // H = S_q R_{q+1} under this shot's projectors
          GemmProblem g;
          memset(&g, 0, sizeof(g));
          g.A = S.d; g.lda = 2 * S.dl; g.B = curR[q + 1]; g.ldb = S.dr; g.C = Hs; g.ldc = 2 * S.dl;
          g.M = 2 * S.dl; g.N = S.dr; g.K = S.dr; g.b_col_stride = 1; g.alpha = 1.0;
          launch_gemm1(g, 0, stream);
          nlaunch += 1;
        }
        const long tot = 2L * S.dl * S.dr;
        launch_dot_batch((const DotProblem*)(ws.base + oP) + 4 * q + (shared_h ? 0 : 2), 2, (int)std::max<long>(1, std::min<long>(DOT_CTAS, tot / 2048)),
                         (double2*)(ws.base + oPart), scal, stream);
        nlaunch += 2;
        CK(cudaMemcpyAsync(h2, scal, 32, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        // a qubit measured twice: its earlier projector is part of w[q]
        const double pb0 = h2[0].real() * w[q][0], pb1 = h2[1].real() * w[q][1];
        const double p0 = std::fabs(pb0) < PROB_EPS ? 0.0 : pb0;
        const double p1 = std::fabs(pb1) < PROB_EPS ? 0.0 : pb1;
        const double r = std::uniform_real_distribution<double>(0.0, 1.0)(rng);
        const int bit = (r <= p0) ? 0 : 1;
        const double pr = bit == 0 ? p0 : p1;
        w[q][bit] *= 1.0 / pr;
        w[q][1 - bit] = 0.0;
        out[(size_t)s * nm + mi] = bit ? '1' : '0';
        // the projector invalidates the conditioned environments on both sides of q; the left one is extended right away from
        // the F just computed:  L_{q+1} = w_b S_b^H F_b
        if (std::isfinite(w[q][bit])) {
          GemmProblem g;
          memset(&g, 0, sizeof(g));
          g.A = S.d + (size_t)bit * S.dl; g.lda = 2 * S.dl; g.B = F + (size_t)bit * S.dl; g.ldb = 2 * S.dl; g.C = env(oLc, q + 1); g.ldc = S.dr;
          g.M = S.dr; g.N = S.dr; g.K = S.dl; g.b_col_stride = 1; g.alpha = w[q][bit];
          launch_gemm1(g, 1, stream);
          nlaunch += 1;
          curL[q + 1] = env(oLc, q + 1);
          Lvalid = q + 1;
        } else Lvalid = std::min(Lvalid, q);
        Rvalid = std::max(Rvalid, q + 1);
      }
    }
    CK(cudaGetLastError());
    return shots;
  }

  // amplitudes with open legs: bits[k] in {0,1,-1}
  void amplitude(int reg, const int8_t* bits, std::vector<cplx>& out) {
    flush();
    const int s0 = reg * nq;
    int nopen = 0;
    for (int k = 0; k < nq; ++k) if (bits[k] < 0) ++nopen;
    if (nopen > 30) throw std::runtime_error("too many open legs");
    // buffer size: rows * max(dr) growing; bound by 2^nopen * max bond * 2
    size_t maxd = 1;
    for (int k = s0; k < s0 + nq; ++k) maxd = std::max(maxd, (size_t)sites[k].dr);
    const size_t elems = ((size_t)1 << nopen) * maxd * 2;
    ws.reset();
    const size_t o0 = ws.reserve(elems * 16), o1 = ws.reserve(elems * 16);
    ensure_ws(ws.off);
    double2* S[2] = {(double2*)(ws.base + o0), (double2*)(ws.base + o1)};
    const cplx one(1, 0);
    CK(cudaMemcpyAsync(S[0], &one, 16, cudaMemcpyHostToDevice, stream));
    int cur = 0;
    size_t rows = 1;
    for (int k = 0; k < nq; ++k) {
      const SiteBuf& sb = sites[s0 + k];
      GemmProblem g;
      memset(&g, 0, sizeof(g));
      g.A = S[cur]; g.lda = (int)rows; g.B = sb.d; g.ldb = sb.dl; g.C = S[cur ^ 1]; g.ldc = (int)rows;
      g.M = (int)rows; g.K = sb.dl; g.alpha = 1.0;
      if (bits[k] < 0) { g.N = 2 * sb.dr; g.b_col_stride = 1; g.b_col_off = 0; }
      else { g.N = sb.dr; g.b_col_stride = 2; g.b_col_off = bits[k]; }
      launch_gemm1(g, 0, stream);
      nlaunch += 1;
      if (bits[k] < 0) rows *= 2;
      cur ^= 1;
    }
    out.resize(rows);
    CK(cudaMemcpyAsync(out.data(), S[cur], 16 * rows, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
  }
};


// =================================================================================== site-sharded group
namespace {

// Contiguous site blocks per device.  by_cost: blocks of equal estimated SVD cost M N min(M, N) on the saturated bond profile
// min(max_bond, 2^k, 2^(n-k)) ("partition by sum chi^3, not by site count"; a gate on bond k is charged to the owner
// of site k, which executes it); otherwise equal counts.  One formula (the reference's two disagree when n % P != 0).
std::vector<std::pair<int, int>> shard_partition(int n, int world, int max_bond, bool by_cost) {
  std::vector<std::pair<int, int>> out;
  if (world < 1 || n < world) throw std::runtime_error("need at least one site per device");
  if (!by_cost || max_bond <= 0 || max_bond >= (1 << 20) || world == 1) {
    int s = 0;
    for (int r = 0; r < world; ++r) {
      const int e = s + n / world + (r < n % world ? 1 : 0);
      out.push_back({s, e});
      s = e;
    }
    return out;
  }
  std::vector<double> dims(n + 1, 1.0), w(n, 0.0);
  for (int k = 0; k + 1 < n; ++k) dims[k + 1] = std::min<double>(max_bond, std::ldexp(1.0, std::min({k + 1, n - 1 - k, 60})));
  double total = 0;
  for (int k = 0; k + 1 < n; ++k) {
    const double m = 2.0 * dims[k], nn = 2.0 * dims[k + 2];
    w[k] = m * nn * std::min(m, nn);
    total += w[k];
  }
  int s = 0;
  double acc = 0;
  for (int r = 0; r < world; ++r) {
    int e;
    if (r == world - 1) e = n;
    else {
      e = s + 1;
      acc += w[s];
      const double target = total * (r + 1) / world;
      while (e < n - (world - 1 - r) && std::fabs(acc + w[e] - target) <= std::fabs(acc - target)) { acc += w[e]; ++e; }
    }
    out.push_back({s, e});
    s = e;
  }
  return out;
}

}  // namespace

namespace {
// The schedule of a flush on a site-sharded group: the gate list (q0, q1 or -1, skip) is cut into dependency layers exactly as
// on one device; per layer every device gets [sites coming home] [boundary sites it lends] [boundary sites it borrows] [its
// gates as ONE batched layer] [borrowed sites handed back].  A gate runs on the owner of its LEFT site.  Every wait (RECV)
// refers to a SEND of an earlier phase of the same layer or of an earlier layer, so the lists cannot deadlock.
std::vector<std::vector<ShardOp>> plan_shard_ops(int ntot, const std::vector<int>& owner, int P, const std::vector<std::array<int, 3>>& gates,
                                                 int* n_slots, int* n_exchanges) {
  std::vector<int> level(ntot, -1);
  std::vector<std::vector<int>> layers;
  for (size_t i = 0; i < gates.size(); ++i) {
    const int q0 = gates[i][0], q1 = gates[i][1];
    if (gates[i][2]) continue;
    int l = level[q0];
    if (q1 >= 0) l = std::max(l, level[q1]);
    ++l;
    if ((int)layers.size() <= l) layers.resize(l + 1);
    layers[l].push_back((int)i);
    level[q0] = l;
    if (q1 >= 0) level[q1] = l;
  }
  std::vector<std::vector<ShardOp>> ops(P);
  std::vector<int> away(ntot, -1);   // site k is on its left neighbour's device; slot of its way back (-1: at home)
  int slots = 0, exch = 0;
  auto op = [](ShardOp::Kind k, int site, int slot) { ShardOp o; o.kind = k; o.site = site; o.slot = slot; return o; };
  for (auto& L : layers) {
    std::vector<std::vector<ShardOp>> back(P), send(P), recv(P), post(P);
    std::vector<ShardOp> run(P);
    for (int d = 0; d < P; ++d) run[d].kind = ShardOp::LAYER;
    for (int gi : L) {
      const int q0 = gates[gi][0], q1 = gates[gi][1];
      const int lo = q1 < 0 ? q0 : std::min(q0, q1), hi = q1 < 0 ? q0 : std::max(q0, q1);
      const int A = owner[lo], B = owner[hi];
      for (int k : {lo, hi})
        if (away[k] >= 0) { back[owner[k]].push_back(op(ShardOp::RECV, k, away[k])); away[k] = -1; }
      if (A != B) {
        const int s1 = slots++, s2 = slots++;
        send[B].push_back(op(ShardOp::SEND, hi, s1));
        recv[A].push_back(op(ShardOp::RECV, hi, s1));
        post[A].push_back(op(ShardOp::SEND, hi, s2));
        away[hi] = s2;
        ++exch;
      }
      run[A].gates.push_back(gi);
    }
    for (int d = 0; d < P; ++d) {
      for (auto* v : {&back[d], &send[d], &recv[d]}) ops[d].insert(ops[d].end(), v->begin(), v->end());
      if (!run[d].gates.empty()) ops[d].push_back(run[d]);
      ops[d].insert(ops[d].end(), post[d].begin(), post[d].end());
    }
  }
  for (int k = 0; k < ntot; ++k)
    if (away[k] >= 0) ops[owner[k]].push_back(op(ShardOp::RECV, k, away[k]));   // every site is home when the flush returns
  *n_slots = slots;
  *n_exchanges = exch;
  return ops;
}
}  // namespace

// Executes the coordinator's queue on the group.  The queue is cut into dependency layers exactly as on one device; a device
// runs, per layer, the gates whose LEFT site it owns as one batched run_layer.  For a gate on a block boundary the right
// owner publishes its boundary site before that layer (SEND), the left owner copies it over NVLink (RECV: one
// cudaMemcpyPeerAsync ordered by an event, no host synchronisation), runs the gate with the rest of its layer and publishes the
// new tensor (SEND); the right owner takes it back only when one of its own later gates touches that site.  All devices
// therefore execute a brickwork layer concurrently; the only waits are on tensors that really cross a boundary.
void mps_b200_handle::group_flush() {
  ShardGroup& G = *grp;
  const int P = (int)G.sub.size();
  // leftover folded 1q gates become queue entries (last in program order)
  for (int q = 0; q < ntot; ++q)
    if (has1q[q]) {
      QGate g;
      g.q0 = q; g.q1 = -1;
      for (int i = 0; i < 4; ++i) g.m[i] = p1q[q][i];
      queue.push_back(g);
      has1q[q] = 0;
    }
  if (queue.empty()) return;
  ++state_ver;
  // per-device op lists (pure host logic: plan_shard_ops, also reachable without a device through mps_shard_plan_debug)
  std::vector<std::array<int, 3>> gl(queue.size());
  for (size_t i = 0; i < queue.size(); ++i) gl[i] = {queue[i].q0, queue[i].q1, queue[i].skip};
  int nslots = 0, nexch = 0;
  std::vector<std::vector<ShardOp>> ops = plan_shard_ops(ntot, G.owner, P, gl, &nslots, &nexch);
  G.slots.clear();
  for (int i = 0; i < nslots; ++i) G.slots.emplace_back();
  G.exchanges += nexch;

  G.abort = false;
  G.first_error.clear();
  auto worker = [&](int d) {
    mps_b200_handle* S = G.sub[d];
    try {
      CK(cudaSetDevice(S->device));
      S->xev_used = 0;
      for (const ShardOp& o : ops[d]) {
        if (G.abort) break;
        if (o.kind == ShardOp::LAYER) {
          S->queue.clear();
          std::vector<int> idx;
          for (int gi : o.gates) { idx.push_back((int)S->queue.size()); S->queue.push_back(queue[gi]); }
          ++S->state_ver;
          S->run_layer(idx);
          S->queue.clear();
        } else if (o.kind == ShardOp::SEND) {
          if (S->xev_used == S->xev.size()) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            S->xev.push_back(e);
          }
          cudaEvent_t e = S->xev[S->xev_used++];
          CK(cudaEventRecord(e, S->stream));
          XferSlot& x = G.slots[o.slot];
          {
            std::lock_guard<std::mutex> lk(x.mu);
            x.ptr = S->sites[o.site].d; x.dl = S->sites[o.site].dl; x.dr = S->sites[o.site].dr; x.src_dev = S->device; x.ev = e;
            x.ready = true;
          }
          x.cv.notify_all();
        } else {
          XferSlot& x = G.slots[o.slot];
          {
            std::unique_lock<std::mutex> lk(x.mu);
            x.cv.wait(lk, [&] { return x.ready || G.abort.load(); });
            if (!x.ready) break;
          }
          // order: wait for the producer, THEN (re)allocate -- the buffer this replaces may still be read by the peer's
          // previous copy, which lies before the producer's event in the peer's stream
          CK(cudaStreamWaitEvent(S->stream, x.ev, 0));
          S->ensure_site(o.site, x.dl, x.dr, false);
          const size_t bytes = (size_t)2 * x.dl * x.dr * sizeof(double2);
          CK(cudaMemcpyPeerAsync(S->sites[o.site].d, S->device, x.ptr, x.src_dev, bytes, S->stream));
          ++S->state_ver;
          std::lock_guard<std::mutex> lk(G.err_mu);
          G.bytes_moved += (double)bytes;
        }
      }
      CK(cudaGetLastError());
    } catch (const std::exception& e) {
      {
        std::lock_guard<std::mutex> lk(G.err_mu);
        if (G.first_error.empty()) G.first_error = "device " + std::to_string(S->device) + ": " + e.what();
      }
      G.abort = true;
      for (auto& x : G.slots) { std::lock_guard<std::mutex> lk(x.mu); x.cv.notify_all(); }
    }
  };
  std::vector<std::thread> th;
  for (int d = 1; d < P; ++d) th.emplace_back(worker, d);
  worker(0);
  for (auto& t : th) t.join();
  queue.clear();
  std::fill(last_touch.begin(), last_touch.end(), -1);
  CK(cudaSetDevice(device));
  if (!G.first_error.empty()) throw std::runtime_error(G.first_error);
}

// ---- observables of a site-sharded group WITHOUT moving the state: every device sweeps its own block, and only the chi x chi
// environment at a block boundary hops to the neighbour (one peer copy ordered by an event).  The left-to-right chain
// runs on the engines' main streams, the right-to-left chain on their second streams, so the two chains overlap.
namespace {

struct GroupEnvs {
  // per device d: environments at the positions k0..k1 of its block (L[k] = everything left of site k, R[k] = everything right of
  // site k-1), optional per-site F_k = L_k S_k and H_k = S_k R_{k+1}, scratch
  struct Dev {
    int k0 = 0, k1 = 0;
    std::vector<double2*> L, R, Fk, Hk;   // indexed by k - k0
    double2 *F = nullptr, *F2 = nullptr, *T = nullptr, *T2 = nullptr, *scal = nullptr, *part = nullptr;
    DotProblem* dots = nullptr;
  };
  std::vector<Dev> dev;
};

// fence_src: the source buffer is scratch that the source stream will overwrite -- make it wait for the copy
void group_peer_copy(mps_b200_handle* dst, cudaStream_t dst_stream, double2* dst_ptr, mps_b200_handle* src, cudaStream_t src_stream, const double2* src_ptr, size_t elems,
                     bool fence_src = false) {
  CK(cudaSetDevice(src->device));
  cudaEvent_t e = src->next_obs_event();
  CK(cudaEventRecord(e, src_stream));
  CK(cudaSetDevice(dst->device));
  CK(cudaStreamWaitEvent(dst_stream, e, 0));
  CK(cudaMemcpyPeerAsync(dst_ptr, dst->device, src_ptr, src->device, elems * sizeof(double2), dst_stream));
  if (fence_src) {
    cudaEvent_t back = dst->next_obs_event();
    CK(cudaEventRecord(back, dst_stream));
    CK(cudaSetDevice(src->device));
    CK(cudaStreamWaitEvent(src_stream, back, 0));
    CK(cudaSetDevice(dst->device));
  }
}

// unconditioned environments of the whole chain; keep_fh also keeps F_k and H_k of every site (for <Z_k>); nscal result slots per device
void group_build_envs(mps_b200_handle* h, GroupEnvs& G, bool keep_fh, int nscal) {
  ShardGroup& grp = *h->grp;
  h->flush();
  const int P = (int)grp.sub.size(), n = h->ntot;
  G.dev.assign(P, GroupEnvs::Dev());
  constexpr int DOT_CTAS = 32;
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev& D = G.dev[d];
    D.k0 = grp.bounds[d].first; D.k1 = grp.bounds[d].second;
    const int nl = D.k1 - D.k0;
    CK(cudaSetDevice(S->device));
    S->ws.reset();
    std::vector<size_t> oL(nl + 1), oR(nl + 1), oF(nl), oH(nl);
    for (int i = 0; i <= nl; ++i) {
      const int k = D.k0 + i;
      const size_t dim = (k == 0) ? 1 : (size_t)grp.sub[grp.owner[k - 1]]->sites[k - 1].dr;
      oL[i] = S->ws.reserve(dim * dim * 16);
      oR[i] = S->ws.reserve(dim * dim * 16);
    }
    size_t ms = 2, me = 1;
    for (int i = 0; i < nl; ++i) {
      const SiteBuf& sb = S->sites[D.k0 + i];
      ms = std::max(ms, (size_t)2 * sb.dl * sb.dr);
      me = std::max({me, (size_t)sb.dr * sb.dr, (size_t)sb.dl * sb.dl});
      if (keep_fh) { oF[i] = S->ws.reserve((size_t)2 * sb.dl * sb.dr * 16); oH[i] = S->ws.reserve((size_t)2 * sb.dl * sb.dr * 16); }
    }
    const size_t oFs = S->ws.reserve(ms * 16), oF2 = S->ws.reserve(ms * 16), oT = S->ws.reserve(me * 16), oT2 = S->ws.reserve(me * 16);
    const size_t oS = S->ws.reserve(16 * (size_t)(nl + 8 + nscal));
    const size_t oP = S->ws.reserve(sizeof(DotProblem) * (size_t)std::max(1, nl));
    const size_t oPart = S->ws.reserve(16 * (size_t)std::max(1, nl) * DOT_CTAS);
    S->ensure_ws(S->ws.off);
    char* b = S->ws.base;
    D.L.resize(nl + 1); D.R.resize(nl + 1);
    for (int i = 0; i <= nl; ++i) { D.L[i] = (double2*)(b + oL[i]); D.R[i] = (double2*)(b + oR[i]); }
    if (keep_fh) {
      D.Fk.resize(nl); D.Hk.resize(nl);
      for (int i = 0; i < nl; ++i) { D.Fk[i] = (double2*)(b + oF[i]); D.Hk[i] = (double2*)(b + oH[i]); }
    }
    D.F = (double2*)(b + oFs); D.F2 = (double2*)(b + oF2); D.T = (double2*)(b + oT); D.T2 = (double2*)(b + oT2);
    D.scal = (double2*)(b + oS); D.dots = (DotProblem*)(b + oP); D.part = (double2*)(b + oPart);
  }
  const cplx one(1, 0);
  // the second streams read the sites the main streams may still be writing (write-back of the last gates): fork first
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    CK(cudaSetDevice(S->device));
    CK(cudaEventRecord(S->fork_ev, S->stream));
    CK(cudaStreamWaitEvent(S->stream2, S->fork_ev, 0));
  }
  // left-to-right chain (main streams)
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev& D = G.dev[d];
    CK(cudaSetDevice(S->device));
    if (d == 0) CK(cudaMemcpyAsync(D.L[0], &one, 16, cudaMemcpyHostToDevice, S->stream));
    else {
      const size_t dim = grp.sub[d - 1]->sites[D.k0 - 1].dr;
      group_peer_copy(S, S->stream, D.L[0], grp.sub[d - 1], grp.sub[d - 1]->stream, G.dev[d - 1].L.back(), dim * dim);
    }
    for (int i = 0; i < D.k1 - D.k0; ++i) S->left_step(S->sites[D.k0 + i], D.L[i], keep_fh ? D.Fk[i] : D.F, D.L[i + 1], 1.0, 1.0);
  }
  // right-to-left chain (second streams)
  for (int d = P - 1; d >= 0; --d) {
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev& D = G.dev[d];
    const int nl = D.k1 - D.k0;
    CK(cudaSetDevice(S->device));
    if (d == P - 1) CK(cudaMemcpyAsync(D.R[nl], &one, 16, cudaMemcpyHostToDevice, S->stream2));
    else {
      const size_t dim = S->sites[D.k1 - 1].dr;
      group_peer_copy(S, S->stream2, D.R[nl], grp.sub[d + 1], grp.sub[d + 1]->stream2, G.dev[d + 1].R[0], dim * dim);
    }
    for (int i = nl - 1; i >= 0; --i) S->right_step(S->sites[D.k0 + i], D.R[i + 1], keep_fh ? D.Hk[i] : D.F2, D.R[i], S->stream2);
    CK(cudaEventRecord(S->join_ev, S->stream2));
    CK(cudaStreamWaitEvent(S->stream, S->join_ev, 0));
  }
  (void)n;
}

// <psi| prod_k diag(w_k) |psi> : one left-to-right chain over the devices
cplx group_sweep_weights(mps_b200_handle* h, const std::vector<std::array<double, 2>>& w) {
  ShardGroup& grp = *h->grp;
  h->flush();
  const int P = (int)grp.sub.size();
  const cplx one(1, 0);
  const double2* prev = nullptr;
  cplx out;
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    const int k0 = grp.bounds[d].first, k1 = grp.bounds[d].second;
    CK(cudaSetDevice(S->device));
    size_t ms = 2, me = 1;
    for (int k = k0; k < k1; ++k) {
      ms = std::max(ms, (size_t)2 * S->sites[k].dl * S->sites[k].dr);
      me = std::max({me, (size_t)S->sites[k].dr * S->sites[k].dr, (size_t)S->sites[k].dl * S->sites[k].dl});
    }
    S->ws.reset();
    const size_t oE0 = S->ws.reserve(me * 16), oE1 = S->ws.reserve(me * 16), oF = S->ws.reserve(ms * 16);
    S->ensure_ws(S->ws.off);
    double2* E[2] = {(double2*)(S->ws.base + oE0), (double2*)(S->ws.base + oE1)};
    double2* F = (double2*)(S->ws.base + oF);
    if (d == 0) CK(cudaMemcpyAsync(E[0], &one, 16, cudaMemcpyHostToDevice, S->stream));
    else {
      const size_t dim = grp.sub[d - 1]->sites[k0 - 1].dr;
      group_peer_copy(S, S->stream, E[0], grp.sub[d - 1], grp.sub[d - 1]->stream, prev, dim * dim);
    }
    int cur = 0;
    for (int k = k0; k < k1; ++k) { S->left_step(S->sites[k], E[cur], F, E[cur ^ 1], w[k][0], w[k][1]); cur ^= 1; }
    prev = E[cur];
    if (d == P - 1) {
      CK(cudaMemcpyAsync(&out, E[cur], 16, cudaMemcpyDeviceToHost, S->stream));
      CK(cudaStreamSynchronize(S->stream));
    }
  }
  CK(cudaSetDevice(h->device));
  return out;
}

// <bits|psi> with open legs on a group: the running (rows x bond) matrix hops from device to device at the block boundaries
// (computeWaveFuncSlice, ExaTnMpsVisitor.cpp:2588-2675; no site leaves its device)
void group_amplitude(mps_b200_handle* h, const int8_t* bits, std::vector<cplx>& out) {
  ShardGroup& grp = *h->grp;
  h->flush();
  const int P = (int)grp.sub.size(), n = h->nq;
  int nopen = 0;
  for (int k = 0; k < n; ++k) if (bits[k] < 0) ++nopen;
  if (nopen > 30) throw std::runtime_error("too many open legs");
  size_t maxd = 1;
  for (int k = 0; k < n; ++k) maxd = std::max(maxd, (size_t)grp.sub[grp.owner[k]]->sites[k].dr);
  const size_t elems = ((size_t)1 << nopen) * maxd * 2;
  const cplx one(1, 0);
  const double2* prev = nullptr;
  size_t rows = 1;
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    const int k0 = grp.bounds[d].first, k1 = grp.bounds[d].second;
    CK(cudaSetDevice(S->device));
    S->ws.reset();
    const size_t o0 = S->ws.reserve(elems * 16), o1 = S->ws.reserve(elems * 16);
    S->ensure_ws(S->ws.off);
    double2* B[2] = {(double2*)(S->ws.base + o0), (double2*)(S->ws.base + o1)};
    if (d == 0) CK(cudaMemcpyAsync(B[0], &one, 16, cudaMemcpyHostToDevice, S->stream));
    else group_peer_copy(S, S->stream, B[0], grp.sub[d - 1], grp.sub[d - 1]->stream, prev, rows * (size_t)grp.sub[d - 1]->sites[k0 - 1].dr);
    int cur = 0;
    for (int k = k0; k < k1; ++k) {
      const SiteBuf& sb = S->sites[k];
      GemmProblem g;
      memset(&g, 0, sizeof(g));
      g.A = B[cur]; g.lda = (int)rows; g.B = sb.d; g.ldb = sb.dl; g.C = B[cur ^ 1]; g.ldc = (int)rows;
      g.M = (int)rows; g.K = sb.dl; g.alpha = 1.0;
      if (bits[k] < 0) { g.N = 2 * sb.dr; g.b_col_stride = 1; g.b_col_off = 0; }
      else { g.N = sb.dr; g.b_col_stride = 2; g.b_col_off = bits[k]; }
      launch_gemm1(g, 0, S->stream);
      S->nlaunch += 1;
      if (bits[k] < 0) rows *= 2;
      cur ^= 1;
    }
    prev = B[cur];
    if (d == P - 1) {
      out.resize(rows);
      CK(cudaMemcpyAsync(out.data(), B[cur], 16 * rows, cudaMemcpyDeviceToHost, S->stream));
      CK(cudaStreamSynchronize(S->stream));
      CK(cudaGetLastError());
    }
  }
  CK(cudaSetDevice(h->device));
}

void group_expval_z_all(mps_b200_handle* h, double* out) {
  ShardGroup& grp = *h->grp;
  GroupEnvs G;
  group_build_envs(h, G, true, 0);
  const int P = (int)grp.sub.size();
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev& D = G.dev[d];
    const int nl = D.k1 - D.k0;
    CK(cudaSetDevice(S->device));
    std::vector<DotProblem> hp(nl);
    for (int i = 0; i < nl; ++i) {
      const SiteBuf& sb = S->sites[D.k0 + i];
      hp[i].F = D.Fk[i]; hp[i].H = D.Hk[i]; hp[i].total = 2L * sb.dl * sb.dr; hp[i].n = sb.dl; hp[i].mode = 0; hp[i].w0 = 1.0; hp[i].w1 = -1.0;
    }
    CK(cudaMemcpyAsync(D.dots, hp.data(), sizeof(DotProblem) * (size_t)nl, cudaMemcpyHostToDevice, S->stream));
    launch_dot_batch(D.dots, nl, 32, D.part, D.scal, S->stream);
    if (d == P - 1) CK(cudaMemcpyAsync(D.scal + nl, D.L[nl], 16, cudaMemcpyDeviceToDevice, S->stream));   // <psi|psi>
  }
  for (int d = 0; d < P; ++d) {
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev& D = G.dev[d];
    const int nl = D.k1 - D.k0;
    CK(cudaSetDevice(S->device));
    std::vector<cplx> hv(nl + 1);
    CK(cudaMemcpyAsync(hv.data(), D.scal, 16 * (size_t)(nl + (d == P - 1 ? 1 : 0)), cudaMemcpyDeviceToHost, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    CK(cudaGetLastError());
    for (int i = 0; i < nl; ++i) out[D.k0 + i] = hv[i].real();
    if (d == P - 1) { h->norm_val[0] = hv[nl].real(); h->norm_ver[0] = h->state_ver; }
  }
  CK(cudaSetDevice(h->device));
}

void group_expval_zz_pairs(mps_b200_handle* h, int np, const int* qi, const int* qj, double* out) {
  ShardGroup& grp = *h->grp;
  for (int t = 0; t < np; ++t)
    if (std::min(qi[t], qj[t]) < 0 || std::max(qi[t], qj[t]) >= h->nq) throw std::runtime_error("qubit index out of range");
  GroupEnvs G;
  group_build_envs(h, G, false, np);
  const int P = (int)grp.sub.size();
  std::vector<int> where(np), slot(np);   // device / result slot of every pair
  std::vector<int> used(P, 0);
  for (int t = 0; t < np; ++t) {
    const int i = std::min(qi[t], qj[t]), j = std::max(qi[t], qj[t]);
    int d = grp.owner[i];
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev* D = &G.dev[d];
    CK(cudaSetDevice(S->device));
    if (i == j) {   // Z^2 = 1: the norm
      mps_b200_handle* SL = grp.sub[P - 1];
      CK(cudaSetDevice(SL->device));
      GroupEnvs::Dev& DL = G.dev[P - 1];
      where[t] = P - 1; slot[t] = used[P - 1]++;
      CK(cudaMemcpyAsync(DL.scal + (DL.k1 - DL.k0) + 8 + slot[t], DL.L.back(), 16, cudaMemcpyDeviceToDevice, SL->stream));
      continue;
    }
    const double2* cur = D->L[i - D->k0];
    double2* bufs[2] = {D->T, D->T2};
    int which = 0;
    for (int k = i; k <= j; ++k) {
      if (grp.owner[k] != d) {   // the running environment hops to the next device
        const int d2 = grp.owner[k];
        mps_b200_handle* S2 = grp.sub[d2];
        GroupEnvs::Dev* D2 = &G.dev[d2];
        const size_t dim = S->sites[k - 1].dr;
        group_peer_copy(S2, S2->stream, D2->T, S, S->stream, cur, dim * dim, true);
        d = d2; S = S2; D = D2;
        cur = D->T; bufs[0] = D->T2; bufs[1] = D->T; which = 0;
      }
      const double wz = (k == i || k == j) ? -1.0 : 1.0;
      S->left_step(S->sites[k], cur, D->F, bufs[which], 1.0, wz);
      cur = bufs[which];
      which ^= 1;
    }
    // close with the right environment behind site j
    where[t] = d; slot[t] = used[d]++;
    const int dlast = S->sites[j].dr;
    launch_trace_pair(cur, D->R[j + 1 - D->k0], dlast, D->scal + (D->k1 - D->k0) + 8 + slot[t], S->stream);
  }
  std::vector<std::vector<cplx>> hv(P);
  for (int d = 0; d < P; ++d) {
    if (!used[d]) continue;
    mps_b200_handle* S = grp.sub[d];
    GroupEnvs::Dev& D = G.dev[d];
    CK(cudaSetDevice(S->device));
    hv[d].resize(used[d]);
    CK(cudaMemcpyAsync(hv[d].data(), D.scal + (D.k1 - D.k0) + 8, 16 * (size_t)used[d], cudaMemcpyDeviceToHost, S->stream));
    CK(cudaStreamSynchronize(S->stream));
    CK(cudaGetLastError());
  }
  for (int t = 0; t < np; ++t) out[t] = hv[where[t]][slot[t]].real();
  CK(cudaSetDevice(h->device));
}

}  // namespace

namespace {
// v1 of the group observables: every site is copied to device 0 (peer copies, ordered by events) and sub[0] evaluates.
mps_b200_handle* group_gather(mps_b200_handle* h) {
  ShardGroup& G = *h->grp;
  h->flush();
  mps_b200_handle* S0 = G.sub[0];
  if (G.gathered_ver == (int)(h->state_ver & 0x7fffffff)) return S0;
  for (size_t d = 1; d < G.sub.size(); ++d) {
    mps_b200_handle* S = G.sub[d];
    CK(cudaSetDevice(S->device));
    if (S->xev.empty()) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); S->xev.push_back(e); }
    CK(cudaEventRecord(S->xev[0], S->stream));
    CK(cudaSetDevice(S0->device));
    CK(cudaStreamWaitEvent(S0->stream, S->xev[0], 0));
    for (int k = G.bounds[d].first; k < G.bounds[d].second; ++k) {
      const SiteBuf& src = S->sites[k];
      S0->ensure_site(k, src.dl, src.dr, false);
      CK(cudaMemcpyPeerAsync(S0->sites[k].d, S0->device, src.d, S->device, (size_t)2 * src.dl * src.dr * sizeof(double2), S0->stream));
    }
  }
  CK(cudaSetDevice(S0->device));
  CK(cudaStreamSynchronize(S0->stream));   // the owners may change their sites as soon as this returns
  ++S0->state_ver;
  G.gathered_ver = (int)(h->state_ver & 0x7fffffff);
  return S0;
}
}  // namespace

// =================================================================================== C ABI
#define API_BEGIN(h)              \
  if (!(h)) return 1;             \
  try {                           \
    CK(cudaSetDevice((h)->device));
#define API_END(h)                \
    return 0;                     \
  } catch (const std::exception& e) { \
    (h)->err = e.what();          \
    return 2;                     \
  } catch (...) {                 \
    (h)->err = "unknown error";   \
    return 3;                     \
  }

extern "C" {

int mps_create(int n_qubits, int n_registers, int max_bond, double svd_cutoff, int gauge, int device, uint64_t seed,
               mps_handle_t* out) {
  if (!out) return 1;
  *out = nullptr;
  mps_b200_handle* h = nullptr;
  try {
    if (getenv("MPS_B200_BACKTRACE")) signal(SIGSEGV, segv_handler);
    if (n_qubits < 1 || n_registers < 1) throw std::runtime_error("n_qubits and n_registers must be >= 1");
    if (gauge < 0 || gauge > 2) throw std::runtime_error("unknown gauge");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw std::runtime_error(std::string("no CUDA device available (this engine has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) throw std::runtime_error("bad device index");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw std::runtime_error("libmps_b200 is built for sm_100a (Blackwell) only");
    h = new mps_b200_handle;
    h->nq = n_qubits; h->nreg = n_registers; h->ntot = n_qubits * n_registers;
    if (max_bond > 0) h->max_bond = max_bond;
    if (svd_cutoff >= 0) h->cutoff = svd_cutoff;
    h->gauge = gauge; h->device = device;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
      cudaMemPool_t pool;
      CK(cudaDeviceGetDefaultMemPool(&pool, device));
      uint64_t keep = UINT64_MAX;   // never hand freed site buffers back to the driver between gates
      CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto& ev : h->ev) CK(cudaEventCreate(&ev));
    for (auto& ev : h->sev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
    {
      // persisting-L2 carve-out for the Jacobi work matrices (see run_layer); both limits are device properties
      h->l2_persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
      h->l2_window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
      if (h->l2_persist_max > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_max) != cudaSuccess) {
        cudaGetLastError();
        h->l2_persist_max = h->l2_window_max = 0;
      }
    }
    h->sites.resize(h->ntot);
    h->norm_ver.assign(h->nreg, 0);
    h->norm_val.assign(h->nreg, 0.0);
    h->has1q.assign(h->ntot, 0);
    h->last_touch.assign(h->ntot, -1);
    h->p1q.resize(h->ntot);
    h->sv.assign(std::max(h->ntot - 1, 1), std::vector<double>{1.0});
    // developer overrides for A/B runs of the whole test-suite
    if (const char* e = getenv("MPS_B200_QR")) h->use_qr = atoi(e) != 0;
    if (const char* e = getenv("MPS_B200_MAX_SWEEPS")) h->max_sweeps = std::max(1, std::min(1000, atoi(e)));
    if (const char* e = getenv("MPS_B200_DBG_MODE")) { jacobi_set_debug_mode(atoi(e)); jacobi_cluster_set_debug(atoi(e)); }
    h->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("MPS_B200_JACOBI_TOL")) h->jacobi_tol = atof(e);
    if (const char* e = getenv("MPS_B200_NULL_TOL")) h->null_tol = atof(e);
    if (const char* e = getenv("MPS_B200_WIDE_TASKS")) h->wide_tasks = atoi(e) != 0;
    if (const char* e = getenv("MPS_B200_CHUNK_MB")) h->chunk_mb = std::max(1, atoi(e));
    if (const char* e = getenv("MPS_B200_L2_PERSIST")) h->l2_persist = atoi(e) != 0;
    if (const char* e = getenv("MPS_B200_JACOBI_CLUSTER")) h->jacobi_cluster = atoi(e) != 0;
    if (const char* e = getenv("MPS_B200_CTAS_PER_SM")) h->ctas_per_sm = std::max(0, std::min(4, atoi(e)));
    if (const char* e = getenv("MPS_B200_SMALL_GEMM")) gemm_set_small_path(atoi(e));
    if (seed) h->rng.seed(seed);
    else { std::random_device rd; h->rng.seed(rd()); }   // RandomEngine.hpp:39-42
    h->reset_state();
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    delete h;
    return 2;
  }
}

int mps_destroy(mps_handle_t h) {
  if (!h) return 0;
  if (h->grp) {
    for (auto* s : h->grp->sub) mps_destroy(s);
    delete h->grp;
    h->grp = nullptr;
  }
  cudaSetDevice(h->device);
  for (auto& e : h->xev) if (e) cudaEventDestroy(e);
  for (auto& e : h->oev) if (e) cudaEventDestroy(e);
  cudaStreamSynchronize(h->stream);
  if (getenv("MPS_B200_DBG_MODE")) { jacobi_print_phase_timing(); jacobi_cluster_print_phase_timing(); }
  for (auto& s : h->sites) if (s.d) cudaFreeAsync(s.d, h->stream);
  for (auto& s : h->snap) if (s.d) cudaFreeAsync(s.d, h->stream);
  cudaStreamSynchronize(h->stream);
  if (h->ws.base) cudaFree(h->ws.base);
  for (int i = 0; i < 2; ++i) if (h->pin[i]) cudaFreeHost(h->pin[i]);
  if (h->pin_rb) cudaFreeHost(h->pin_rb);
  for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->sev) if (ev) cudaEventDestroy(ev);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  if (h->join_ev) cudaEventDestroy(h->join_ev);
  if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}


int mps_create_sharded(int n_qubits, int max_bond, double svd_cutoff, int gauge, int n_devices, const int* devices, int partition_by_cost,
                       uint64_t seed, mps_handle_t* out) {
  if (!out) return 1;
  *out = nullptr;
  if (n_devices < 1 || !devices) { g_create_error = "mps_create_sharded: empty device list"; return 2; }
  mps_handle_t h = nullptr;
  int rc = mps_create(n_qubits, 1, max_bond, svd_cutoff, gauge, devices[0], seed, &h);
  if (rc) return rc;
  if (n_devices == 1) { *out = h; return 0; }
  try {
    if (n_qubits < n_devices) throw std::runtime_error("mps_create_sharded: need at least one site per device");
    ShardGroup* G = new ShardGroup;
    h->grp = G;
    G->bounds = shard_partition(n_qubits, n_devices, max_bond, partition_by_cost != 0);
    G->owner.resize(n_qubits);
    for (int d = 0; d < n_devices; ++d)
      for (int k = G->bounds[d].first; k < G->bounds[d].second; ++k) G->owner[k] = d;
    for (int d = 0; d < n_devices; ++d) {
      mps_handle_t s = nullptr;
      if (mps_create(n_qubits, 1, max_bond, svd_cutoff, gauge, devices[d], seed, &s)) throw std::runtime_error("device " + std::to_string(devices[d]) + ": " + g_create_error);
      G->sub.push_back(s);
    }
    // direct NVLink paths between neighbouring devices: peer access for plain allocations and for the stream-ordered pool
    for (int d = 0; d < n_devices; ++d)
      for (int e : {d - 1, d + 1}) {
        if (e < 0 || e >= n_devices || devices[e] == devices[d]) continue;   // a device may hold several blocks (single-GPU tests)
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, devices[d], devices[e]));
        if (!can) continue;
        CK(cudaSetDevice(devices[d]));
        cudaError_t pe = cudaDeviceEnablePeerAccess(devices[e], 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CK(pe);
        cudaGetLastError();
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, devices[e]));
        cudaMemAccessDesc desc;
        memset(&desc, 0, sizeof(desc));
        desc.location.type = cudaMemLocationTypeDevice;
        desc.location.id = devices[d];
        desc.flags = cudaMemAccessFlagsProtReadWrite;
        CK(cudaMemPoolSetAccess(pool, &desc, 1));
      }
    CK(cudaSetDevice(devices[0]));
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    mps_destroy(h);
    return 2;
  }
}

int mps_shard_partition(int n_qubits, int n_devices, int max_bond, int partition_by_cost, int* first_site) {
  try {
    const auto b = shard_partition(n_qubits, n_devices, max_bond, partition_by_cost != 0);
    for (int d = 0; d < n_devices; ++d) first_site[d] = b[d].first;
    first_site[n_devices] = n_qubits;
    return 0;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return 2;
  }
}

int mps_shard_plan_debug(int n_qubits, int n_devices, const int* first_site, int count, const int* q0, const int* q1, int* out, int cap, int* used) {
  try {
    std::vector<int> owner(n_qubits, 0);
    for (int d = 0; d < n_devices; ++d)
      for (int k = first_site[d]; k < first_site[d + 1]; ++k) owner[k] = d;
    std::vector<std::array<int, 3>> gl(count);
    for (int i = 0; i < count; ++i) gl[i] = {q0[i], q1[i], 0};
    int ns = 0, ne = 0;
    const auto ops = plan_shard_ops(n_qubits, owner, n_devices, gl, &ns, &ne);
    int n = 0;
    auto put = [&](int v) { if (n < cap) out[n] = v; ++n; };
    for (int d = 0; d < n_devices; ++d)
      for (const ShardOp& o : ops[d]) {   // record: device, kind (0 LAYER, 1 SEND, 2 RECV), site, slot, number of gates, gate indices
        put(d); put((int)o.kind); put(o.site); put(o.slot); put((int)o.gates.size());
        for (int g : o.gates) put(g);
      }
    *used = n;
    return n <= cap ? 0 : 2;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return 2;
  }
}

int mps_shard_layout(mps_handle_t h, int* n_devices, int* first_site /* n_devices + 1 entries, may be NULL */) {
  API_BEGIN(h)
  const int P = h->grp ? (int)h->grp->sub.size() : 1;
  if (n_devices) *n_devices = P;
  if (first_site) {
    for (int d = 0; d < P; ++d) first_site[d] = h->grp ? h->grp->bounds[d].first : 0;
    first_site[P] = h->ntot;
  }
  API_END(h)
}

const char* mps_last_error(mps_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mps_reset(mps_handle_t h) {
  API_BEGIN(h)
  if (h->grp) {
    h->queue.clear();
    std::fill(h->has1q.begin(), h->has1q.end(), 0);
    std::fill(h->last_touch.begin(), h->last_touch.end(), -1);
    ++h->state_ver;
    for (auto* s : h->grp->sub) { CK(cudaSetDevice(s->device)); s->reset_state(); }
    CK(cudaSetDevice(h->device));
  } else h->reset_state();
  API_END(h)
}
int mps_snapshot(mps_handle_t h) {
  API_BEGIN(h)
  if (h->grp) {
    h->flush();
    for (auto* s : h->grp->sub) { CK(cudaSetDevice(s->device)); s->snapshot_state(); }
    CK(cudaSetDevice(h->device));
    h->has_snap = true;
  } else h->snapshot_state();
  API_END(h)
}
int mps_restore(mps_handle_t h) {
  API_BEGIN(h)
  if (h->grp) {
    if (!h->has_snap) throw std::runtime_error("mps_restore without mps_snapshot");
    h->flush();
    ++h->state_ver;
    for (auto* s : h->grp->sub) { CK(cudaSetDevice(s->device)); s->restore_state(); }
    CK(cudaSetDevice(h->device));
  } else h->restore_state();
  API_END(h)
}

int mps_set_option(mps_handle_t h, const char* key, double value) {
  API_BEGIN(h)
  std::string k(key);
  if (h->grp) {   // the coordinator keeps the queue-side options (fuse_1q, fuse_2q, layer_batch); the engines get everything
    h->flush();
    for (auto* s : h->grp->sub)
      if (mps_set_option(s, key, value)) throw std::runtime_error(s->err);
    CK(cudaSetDevice(h->device));
  }
  // options that change how queued gates would be executed flush the queue first
  if (k == "cutoff_on_sqrt") { h->flush(); h->cutoff_on_sqrt = value != 0; }
  else if (k == "fuse_1q") { h->flush(); h->fuse_1q = value != 0; }
  else if (k == "fuse_2q") { h->flush(); h->fuse_2q = value != 0; }
  else if (k == "renormalize") { h->flush(); h->renorm = value != 0; }
  else if (k == "jacobi_tol") { h->flush(); h->jacobi_tol = value; }
  else if (k == "null_tol") { h->flush(); h->null_tol = value; }
  else if (k == "jacobi_max_sweeps") { h->flush(); h->max_sweeps = std::max(1, std::min(1000, (int)value)); }
  else if (k == "profile") h->profile = value != 0;
  else if (k == "layer_batch") { h->flush(); h->layer_batch = value != 0; }
  else if (k == "qr_prereduce") { h->flush(); h->use_qr = value != 0; }
  else if (k == "jacobi_wide_tasks") { h->flush(); h->wide_tasks = value != 0; }
  else if (k == "jacobi_ctas_per_sm") { h->flush(); h->ctas_per_sm = std::max(0, std::min(4, (int)value)); }
  else if (k == "jacobi_chunk_mb") { h->flush(); h->chunk_mb = std::max(1, (int)value); }
  else if (k == "l2_persist") { h->flush(); h->l2_persist = value != 0; }
  else if (k == "jacobi_cluster") { h->flush(); h->jacobi_cluster = value != 0; }
  else if (k == "norm_guard") { h->flush(); h->norm_guard = std::max(0, std::min(2, (int)value)); }
  else if (k == "max_bond") { h->flush(); h->max_bond = value > 0 ? (int)value : INT_MAX - 1; }
  else if (k == "svd_cutoff") { h->flush(); h->cutoff = value >= 0 ? value : DBL_MIN; }
  else if (k == "gauge") { h->flush(); if (value < 0 || value > 2) throw std::runtime_error("unknown gauge"); h->gauge = (int)value; }
  else throw std::runtime_error("unknown option: " + k);
  API_END(h)
}

int mps_apply_1q(mps_handle_t h, int q, const double m[8]) {
  API_BEGIN(h) h->push_1q(q, reinterpret_cast<const cplx*>(m)); API_END(h)
}
int mps_apply_2q(mps_handle_t h, int q0, int q1, const double m[32]) {
  API_BEGIN(h) h->push_2q(q0, q1, reinterpret_cast<const cplx*>(m)); API_END(h)
}
int mps_apply_layer(mps_handle_t h, int count, const int* q0, const int* q1, const double* mats) {
  API_BEGIN(h)
  for (int i = 0; i < count; ++i) h->push_2q(q0[i], q1[i], reinterpret_cast<const cplx*>(mats + 32 * (size_t)i));
  API_END(h)
}
int mps_apply_gates(mps_handle_t h, int count, const int* q0, const int* q1, const double* mats) {
  API_BEGIN(h)
  for (int i = 0; i < count; ++i) {
    const cplx* m = reinterpret_cast<const cplx*>(mats + 32 * (size_t)i);
    if (q1[i] < 0) h->push_1q(q0[i], m);
    else h->push_2q(q0[i], q1[i], m);
  }
  API_END(h)
}
int mps_flush(mps_handle_t h) { API_BEGIN(h) h->flush(); API_END(h) }
int mps_sync(mps_handle_t h) {
  API_BEGIN(h)
  h->flush();
  if (h->grp)
    for (auto* s : h->grp->sub) { CK(cudaSetDevice(s->device)); CK(cudaStreamSynchronize(s->stream)); }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  API_END(h)
}

int mps_norm(mps_handle_t h, int reg, double* out) {
  API_BEGIN(h)
  if (reg < 0 || reg >= h->nreg) throw std::runtime_error("bad register");
  h->flush();
  if (h->norm_ver[reg] != h->state_ver) {
    std::vector<std::array<double, 2>> w(h->nq, {1.0, 1.0});
    h->norm_val[reg] = (h->grp ? group_sweep_weights(h, w) : h->sweep_weights(reg, w)).real();
    h->norm_ver[reg] = h->state_ver;
  }
  *out = h->norm_val[reg];
  API_END(h)
}
int mps_expval_z(mps_handle_t h, int reg, int nq, const int* qubits, double* out) {
  API_BEGIN(h)
  if (reg < 0 || reg >= h->nreg) throw std::runtime_error("bad register");
  std::vector<std::array<double, 2>> w(h->nq, {1.0, 1.0});
  for (int i = 0; i < nq; ++i) {
    if (qubits[i] < 0 || qubits[i] >= h->nq) throw std::runtime_error("qubit index out of range");
    w[qubits[i]][1] = -w[qubits[i]][1];
  }
  *out = (h->grp ? group_sweep_weights(h, w) : h->sweep_weights(reg, w)).real();
  API_END(h)
}
int mps_expval_z_all(mps_handle_t h, int reg, double* out_n) {
  API_BEGIN(h)
  if (reg < 0 || reg >= h->nreg) throw std::runtime_error("bad register");
  if (h->grp) group_expval_z_all(h, out_n);
  else h->expval_z_all(reg, out_n);
  API_END(h)
}
int mps_expval_zz_pairs(mps_handle_t h, int reg, int npairs, const int* qi, const int* qj, double* out) {
  API_BEGIN(h)
  if (reg < 0 || reg >= h->nreg) throw std::runtime_error("bad register");
  if (h->grp) group_expval_zz_pairs(h, npairs, qi, qj, out);
  else h->expval_zz_pairs(reg, npairs, qi, qj, out);
  API_END(h)
}
int mps_amplitude(mps_handle_t h, int reg, const int8_t* bits, double* out, size_t* len) {
  API_BEGIN(h)
  if (reg < 0 || reg >= h->nreg) throw std::runtime_error("bad register");
  std::vector<cplx> v;
  if (h->grp) group_amplitude(h, bits, v);
  else h->amplitude(reg, bits, v);
  if (out) memcpy(out, v.data(), 16 * v.size());
  if (len) *len = v.size();
  API_END(h)
}
int mps_statevector(mps_handle_t h, int reg, double* out) {
  API_BEGIN(h)
  if (reg < 0 || reg >= h->nreg) throw std::runtime_error("bad register");
  if (h->nq > 30) throw std::runtime_error("state vector limited to 30 qubits");
  std::vector<int8_t> bits(h->nq, -1);
  std::vector<cplx> v;
  if (h->grp) group_amplitude(h, bits.data(), v);
  else h->amplitude(reg, bits.data(), v);
  memcpy(out, v.data(), 16 * v.size());
  API_END(h)
}

int mps_measure(mps_handle_t h, int q) {
  API_BEGIN(h)
  mps_b200_handle* e = h->grp ? h->grp->sub[0] : h;
  if (q < 0 || q >= e->nq) throw std::runtime_error("qubit index out of range");
  e->measure.push_back(q);
  API_END(h)
}
int mps_clear_measure(mps_handle_t h) { API_BEGIN(h) (h->grp ? h->grp->sub[0] : h)->measure.clear(); API_END(h) }
int mps_seed(mps_handle_t h, uint64_t seed) { API_BEGIN(h) (h->grp ? h->grp->sub[0] : h)->rng.seed(seed); API_END(h) }

int mps_n_measured(mps_handle_t h, int* out) { API_BEGIN(h)
  mps_b200_handle* e = h->grp ? h->grp->sub[0] : h; *out = (int)e->measure.size(); API_END(h) }

int mps_sample(mps_handle_t h, int reg, int shots, char* out, size_t out_cap, int* n_out) {
  API_BEGIN(h)
  mps_b200_handle* e = h->grp ? group_gather(h) : h;   // site-sharded group: evaluated by the engine of device 0
  if (reg < 0 || reg >= e->nreg) throw std::runtime_error("bad register");
  const int nm = (int)e->measure.size();
  if (shots > 0 && (size_t)shots * (size_t)nm > out_cap)
    throw std::runtime_error("mps_sample: output buffer too small (" + std::to_string(out_cap) + " chars for " + std::to_string(shots) + " shots x " +
                             std::to_string(nm) + " measured qubits; see mps_n_measured)");
  int produced = 0;
  if (nm > 0 && shots > 0) {
    if (e->nq < 20) {   // MAX_NUMBER_QUBITS_FOR_STATE_VEC, ExaTnMpsVisitor.cpp:55
      // GenerateSamples (GateMatrixAlgebra.hpp:125-156): draw, sort, walk the CDF in state-index order
      std::vector<int8_t> bits(e->nq, -1);
      std::vector<cplx> sv;
      e->amplitude(reg, bits.data(), sv);
      std::vector<double> rs;
      rs.reserve(shots + 1);
      for (int i = 0; i < shots; ++i) rs.push_back(std::uniform_real_distribution<double>(0.0, 1.0)(e->rng));
      std::sort(rs.begin(), rs.end());
      double csum = 0.0;
      size_t m = 0;
      for (size_t k = 0; k < sv.size(); ++k) {
        csum += std::norm(sv[k]);
        while (m < (size_t)shots && rs[m] < csum) {
          for (int i = 0; i < nm; ++i) out[m * nm + i] = (k & (1ULL << e->measure[i])) ? '1' : '0';
          ++m;
        }
      }
      produced = (int)m;
    } else {
      produced = e->sample_rdm(reg, shots, out);
    }
  }
  if (n_out) *n_out = produced;
  API_END(h)
}

int mps_bond_dims(mps_handle_t h, int* out) {
  API_BEGIN(h)
  h->flush();
  for (int k = 0; k + 1 < h->ntot; ++k) out[k] = (h->grp ? h->grp->sub[h->grp->owner[k]] : h)->sites[k].dr;
  API_END(h)
}
int mps_singular_values(mps_handle_t h, int bond, double* out, int cap, int* count) {
  API_BEGIN(h)
  h->flush();
  if (bond < 0 || bond >= h->ntot - 1) throw std::runtime_error("bad bond index");
  const auto& v = (h->grp ? h->grp->sub[h->grp->owner[bond]] : h)->sv[bond];   // a bond's gates run on the owner of its left site
  const int c = std::min<int>(cap, (int)v.size());
  for (int i = 0; i < c; ++i) out[i] = v[i];
  if (count) *count = (int)v.size();
  API_END(h)
}
int mps_discarded_weight(mps_handle_t h, double* out) {
  API_BEGIN(h)
  h->flush();
  *out = h->discarded;
  if (h->grp) for (auto* s : h->grp->sub) *out += s->discarded;
  API_END(h)
}
int mps_fidelity_estimate(mps_handle_t h, double* out) {
  API_BEGIN(h)
  h->flush();
  double lf = h->log_fidelity;
  if (h->grp) for (auto* s : h->grp->sub) lf += s->log_fidelity;
  *out = std::exp(lf);
  API_END(h)
}
// a site of a group lives in the engine of its owner device
static mps_b200_handle* site_engine(mps_handle_t h, int k) {
  if (k < 0 || k >= h->ntot) throw std::runtime_error("bad site index");
  if (!h->grp) return h;
  mps_b200_handle* e = h->grp->sub[h->grp->owner[k]];
  CK(cudaSetDevice(e->device));
  return e;
}
int mps_get_site(mps_handle_t h, int k, double* out, int shape[3]) {
  API_BEGIN(h)
  h->flush();
  mps_b200_handle* e = site_engine(h, k);
  const SiteBuf& s = e->sites[k];
  shape[0] = s.dl; shape[1] = 2; shape[2] = s.dr;
  if (out) {
    CK(cudaMemcpyAsync(out, s.d, (size_t)2 * s.dl * s.dr * 16, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  API_END(h)
}
int mps_set_site(mps_handle_t h, int k, const double* in, int dl, int dr) {
  API_BEGIN(h)
  h->flush();
  ++h->state_ver;   // the caller may change the tensor
  mps_b200_handle* e = site_engine(h, k);
  ++e->state_ver;
  e->ensure_site(k, dl, dr, false);
  CK(cudaMemcpyAsync(e->sites[k].d, in, (size_t)2 * dl * dr * 16, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  API_END(h)
}
// all the sites of a list in one call: the copies are queued back to back and waited for ONCE (a per-site call costs a stream
// synchronisation each; a host-resident state of 50 sites is 100 of them per round trip)
int mps_set_sites(mps_handle_t h, int count, const int* k, const double* const* in, const int* dl, const int* dr) {
  API_BEGIN(h)
  h->flush();
  ++h->state_ver;
  std::vector<mps_b200_handle*> touched;
  for (int i = 0; i < count; ++i) {
    mps_b200_handle* e = site_engine(h, k[i]);
    ++e->state_ver;
    e->ensure_site(k[i], dl[i], dr[i], false);
    CK(cudaMemcpyAsync(e->sites[k[i]].d, in[i], (size_t)2 * dl[i] * dr[i] * 16, cudaMemcpyHostToDevice, e->stream));
    if (std::find(touched.begin(), touched.end(), e) == touched.end()) touched.push_back(e);
  }
  for (auto* e : touched) { CK(cudaSetDevice(e->device)); CK(cudaStreamSynchronize(e->stream)); }
  API_END(h)
}
int mps_get_sites(mps_handle_t h, int count, const int* k, double* const* out) {
  API_BEGIN(h)
  h->flush();
  std::vector<mps_b200_handle*> touched;
  for (int i = 0; i < count; ++i) {
    mps_b200_handle* e = site_engine(h, k[i]);
    const SiteBuf& s = e->sites[k[i]];
    CK(cudaMemcpyAsync(out[i], s.d, (size_t)2 * s.dl * s.dr * 16, cudaMemcpyDeviceToHost, e->stream));
    if (std::find(touched.begin(), touched.end(), e) == touched.end()) touched.push_back(e);
  }
  for (auto* e : touched) { CK(cudaSetDevice(e->device)); CK(cudaStreamSynchronize(e->stream)); }
  API_END(h)
}
int mps_site_device_ptr(mps_handle_t h, int k, void** dptr, int shape[3]) {
  API_BEGIN(h)
  h->flush();
  ++h->state_ver;   // the caller may change the tensor
  mps_b200_handle* e = site_engine(h, k);
  ++e->state_ver;
  CK(cudaStreamSynchronize(e->stream));
  const SiteBuf& s = e->sites[k];
  *dptr = s.d;
  shape[0] = s.dl; shape[1] = 2; shape[2] = s.dr;
  API_END(h)
}
int mps_resize_site(mps_handle_t h, int k, int dl, int dr, void** dptr) {
  API_BEGIN(h)
  h->flush();
  ++h->state_ver;   // the caller may change the tensor
  mps_b200_handle* e = site_engine(h, k);
  ++e->state_ver;
  e->ensure_site(k, dl, dr, false);
  CK(cudaStreamSynchronize(e->stream));
  *dptr = e->sites[k].d;
  API_END(h)
}
int mps_stats(mps_handle_t h, double* out, int cap) {
  API_BEGIN(h)
  h->flush();
  CK(cudaStreamSynchronize(h->stream));
  double v[15] = {h->n2q, h->n1q, h->nlayers, h->nsweeps, h->nlaunch, h->ms_theta, h->ms_svd, h->ms_wb, h->ms_qr, 0.0, h->nfused2q,
                  h->nonconverged, h->guard_violations, 0.0, 0.0};
  if (h->grp) {
    for (auto* s : h->grp->sub) {
      CK(cudaSetDevice(s->device));
      CK(cudaStreamSynchronize(s->stream));
      const double w[13] = {s->n2q, s->n1q, s->nlayers, s->nsweeps, s->nlaunch, s->ms_theta, s->ms_svd, s->ms_wb, s->ms_qr, jacobi_dmma_flops() + jacobi_cluster_dmma_flops(), 0.0,
                            s->nonconverged, s->guard_violations};
      for (int i = 0; i < 13; ++i) v[i] += w[i];
    }
    v[13] = h->grp->exchanges;
    v[14] = h->grp->bytes_moved;
    CK(cudaSetDevice(h->device));
  } else v[9] = jacobi_dmma_flops() + jacobi_cluster_dmma_flops();
  for (int i = 0; i < cap && i < 15; ++i) out[i] = v[i];
  API_END(h)
}

int mps_get_stream(mps_handle_t h, void** stream) {
  API_BEGIN(h)
  *stream = (void*)h->stream;
  API_END(h)
}

}  // extern "C"
