// See B200MpsVisitor.hpp.  Reference behaviour followed line by line (nothing copied):
//   initialize  ExaTnMpsVisitor.cpp:173-346   options, reset to |0...0>
//   visit(*)    ExaTnMpsVisitor.cpp:870-1085  gate -> applyGate; Swap pre-sorts its bits (:1030-1033); Measure records (:991-994)
//   finalize    ExaTnMpsVisitor.cpp:576-670   "norm", "exp-val-z" (shots < 1) or bit strings (shots >= 1)
#include "B200MpsVisitor.hpp"

#include <chrono>

#include <algorithm>
#include <cmath>

namespace tnqvm {

int b200GateMatrix(const std::string& name, const std::vector<double>& p, std::complex<double> m[16]) {
  typedef std::complex<double> C;
  const C I(0.0, 1.0);
  auto P = [&](size_t i) { return i < p.size() ? p[i] : 0.0; };
  auto set2 = [&](C a, C b, C c, C d) { m[0] = a; m[1] = b; m[2] = c; m[3] = d; return 2; };
  auto ctrl = [&](C a, C b, C c, C d) {   // controlled-U on (q0 = control, q1 = target), index 2*b0 + b1
    for (int i = 0; i < 16; ++i) m[i] = 0.0;
    m[0] = m[5] = 1.0;
    m[10] = a; m[11] = b; m[14] = c; m[15] = d;
    return 4;
  };
  const double s2 = std::sqrt(0.5);
  if (name == "H") return set2(s2, s2, s2, -s2);
  if (name == "X") return set2(0, 1, 1, 0);
  if (name == "Y") return set2(0, -I, I, 0);
  if (name == "Z") return set2(1, 0, 0, -1);
  if (name == "T") return set2(1, 0, 0, std::exp(I * (M_PI / 4)));
  if (name == "Tdg") return set2(1, 0, 0, std::exp(-I * (M_PI / 4)));
  if (name == "S") return set2(1, 0, 0, I);
  if (name == "Sdg") return set2(1, 0, 0, -I);
  if (name == "Rx") { const double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2); return set2(c, -I * s, -I * s, c); }
  if (name == "Ry") { const double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2); return set2(c, -s, s, c); }
  if (name == "Rz") return set2(std::exp(-I * (P(0) / 2)), 0, 0, std::exp(I * (P(0) / 2)));
  if (name == "U" || name == "U3") {
    const double c = std::cos(P(0) / 2), s = std::sin(P(0) / 2);
    return set2(c, -std::exp(I * P(2)) * s, std::exp(I * P(1)) * s, std::exp(I * (P(1) + P(2))) * c);
  }
  if (name == "CNOT" || name == "CX") return ctrl(0, 1, 1, 0);
  if (name == "CZ") return ctrl(1, 0, 0, -1);
  if (name == "CY") return ctrl(0, -I, I, 0);
  if (name == "CH") return ctrl(s2, s2, s2, -s2);
  if (name == "CRZ") return ctrl(std::exp(-I * (P(0) / 2)), 0, 0, std::exp(I * (P(0) / 2)));
  if (name == "CPhase") return ctrl(1, 0, 0, std::exp(I * P(0)));
  if (name == "Swap" || name == "iSwap" || name == "fSim") {
    for (int i = 0; i < 16; ++i) m[i] = 0.0;
    m[0] = 1.0;
    if (name == "Swap") { m[6] = m[9] = 1.0; m[15] = 1.0; }
    else if (name == "iSwap") { m[6] = m[9] = I; m[15] = 1.0; }
    else {
      const double c = std::cos(P(0)), s = std::sin(P(0));
      m[5] = m[10] = c; m[6] = m[9] = -I * s; m[15] = std::exp(-I * P(1));
    }
    return 4;
  }
  return set2(1, 0, 0, 1);   // unknown name: identity (ExatnUtils.cpp:112)
}

B200MpsVisitor::B200MpsVisitor() {}

B200MpsVisitor::~B200MpsVisitor() {
  if (m_handle) mps_destroy(m_handle);
}

void B200MpsVisitor::check(int rc, const char* what) const {
  if (rc != 0) {
    const char* msg = mps_last_error(m_handle);
    xacc::error(std::string("B200MpsVisitor: ") + what + " failed: " + (msg ? msg : "?"));
  }
}

void B200MpsVisitor::CallStat::add(double s) {
  if (calls == 0) { mx = s; mn = s; }
  ++calls; total += s; mx = std::max(mx, s); mn = std::min(mn, s);
}
namespace {
struct StatTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double secs() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
}  // namespace

void B200MpsVisitor::initialize(std::shared_ptr<AcceleratorBuffer> in_buffer, int nbShots) {
  StatTimer timer;
  m_statInit = m_statFinalize = m_stat1q = m_stat2q = CallStat();
  double svdCutoff = -1.0;   // -> DBL_MIN inside the engine (ExaTnMpsVisitor.cpp:257)
  if (options.keyExists<double>("svd-cutoff")) svdCutoff = options.get<double>("svd-cutoff");
  int maxBondDim = 0;        // -> unlimited (ExaTnMpsVisitor.cpp:265)
  if (options.keyExists<int>("max-bond-dim")) maxBondDim = options.get<int>("max-bond-dim");
  int device = 0, gauge = MPS_GAUGE_REFERENCE;
  if (options.keyExists<int>("b200-device")) device = options.get<int>("b200-device");
  if (options.keyExists<int>("b200-gauge")) gauge = options.get<int>("b200-gauge");
  uint64_t seed = 0;
  if (options.keyExists<int>("seed")) seed = (uint64_t)options.get<int>("seed");   // TNQVM.hpp:114-117

  m_buffer = std::move(in_buffer);
  buffer = m_buffer;
  m_measureQubits.clear();
  m_shotCount = nbShots;
  const int n = m_buffer->size();
  // the registered service instance is reused across execute() calls (TNQVM.cpp:109): a fresh state every time.  The engine
  // handle (device workspace, pinned staging, site buffers) is kept when the register shape and the device placement are the
  // same, and only put back to |0...0> with this call's options -- re-creating it costs more than a small circuit.
  std::vector<int> devices;
  if (options.keyExists<std::vector<int>>("b200-devices")) devices = options.get<std::vector<int>>("b200-devices");
  const bool byCount = options.keyExists<bool>("b200-partition-by-count") && options.get<bool>("b200-partition-by-count");
  if (devices.empty()) devices.push_back(device);
  const bool reuse = m_handle && m_nQubits == n && m_devices == devices && m_byCount == byCount && m_maxBondAtCreate == maxBondDim;
  if (m_handle && !reuse) { mps_destroy(m_handle); m_handle = nullptr; }
  m_nQubits = n;
  if (reuse) {
    check(mps_reset(m_handle), "reset");
    check(mps_set_option(m_handle, "svd_cutoff", svdCutoff), "set_option");
    check(mps_set_option(m_handle, "gauge", (double)gauge), "set_option");
    check(mps_set_option(m_handle, "cutoff_on_sqrt", 0.0), "set_option");
    check(mps_set_option(m_handle, "fuse_2q", 0.0), "set_option");
    check(mps_set_option(m_handle, "profile", 0.0), "set_option");
    if (seed) check(mps_seed(m_handle, seed), "seed");
  } else {
    // {"b200-devices", vector<int>}: shard the sites over these GPUs of the box (what the MPI build of the reference does with one
    // rank per site block, ExaTnMpsVisitor.cpp:347-531); {"b200-partition-by-count", true} = equal site counts instead of equal cost
    const int rc = devices.size() > 1
                       ? mps_create_sharded(n, maxBondDim, svdCutoff, gauge, (int)devices.size(), devices.data(), byCount ? 0 : 1, seed, &m_handle)
                       : mps_create(n, 1, maxBondDim, svdCutoff, gauge, devices[0], seed, &m_handle);
    if (rc != 0) {
      const char* msg = mps_last_error(nullptr);
      xacc::error(std::string("B200MpsVisitor: cannot create the MPS engine: ") + (msg ? msg : "?"));
    }
    m_devices = devices; m_byCount = byCount; m_maxBondAtCreate = maxBondDim;
  }
  if (options.keyExists<bool>("b200-cutoff-on-sqrt") && options.get<bool>("b200-cutoff-on-sqrt"))
    check(mps_set_option(m_handle, "cutoff_on_sqrt", 1.0), "set_option");
  // not a reference option: merge consecutive 2q gates on one site pair into one 4x4 (CX.Rz.CX, Swap.Swap) before the GPU
  if (options.keyExists<bool>("b200-fuse-2q") && options.get<bool>("b200-fuse-2q"))
    check(mps_set_option(m_handle, "fuse_2q", 1.0), "set_option");
  // per-phase GPU timings (CUDA events around the merge GEMM / SVD / truncate+write-back of every layer) for getExecutionInfo()
  if (options.keyExists<bool>("b200-profile") && options.get<bool>("b200-profile"))
    check(mps_set_option(m_handle, "profile", 1.0), "set_option");
  m_statsAtInit.assign(13, 0.0);
  check(mps_stats(m_handle, m_statsAtInit.data(), 13), "stats");
  m_statInit.add(timer.secs());
}

void B200MpsVisitor::applyGate(xacc::Instruction& inst) {
  std::vector<double> params;
  for (int i = 0; i < (int)inst.getParameters().size(); ++i) params.push_back(inst.getParameter(i).as<double>());
  std::complex<double> m[16];
  const int dim = b200GateMatrix(inst.name(), params, m);
  const auto bits = inst.bits();
  StatTimer timer;
  if (dim == 2) {
    check(mps_apply_1q(m_handle, (int)bits[0], reinterpret_cast<const double*>(m)), "apply_1q");
    m_stat1q.add(timer.secs());
  } else {
    if (bits.size() != 2) xacc::error("two-qubit gate with " + std::to_string(bits.size()) + " bits");
    check(mps_apply_2q(m_handle, (int)bits[0], (int)bits[1], reinterpret_cast<const double*>(m)), "apply_2q");
    m_stat2q.add(timer.secs());
  }
}

void B200MpsVisitor::visit(Swap& g) {
  // ExaTnMpsVisitor.cpp:1030-1033: bits sorted descending before the gate tensor is applied (Swap is symmetric)
  auto b = g.bits();
  if (b.size() == 2 && b[0] < b[1]) g.setBits({b[1], b[0]});
  applyGate(g);
}

void B200MpsVisitor::visit(Measure& g) {
  m_measureQubits.push_back(g.bits()[0]);
  check(mps_measure(m_handle, (int)g.bits()[0]), "measure");
}

void B200MpsVisitor::exportStats() {
  // Buckets the host can time (gates are queued asynchronously, so the per-call host time of a gate is its queueing cost; the
  // GPU time of the three phases of the 2q step comes from the engine's own events when "b200-profile" is set).
  auto put = [&](const std::string& name, const CallStat& c) {
    executionInfo.insert(name + " [calls]", c.calls);
    executionInfo.insert(name + " [secs]", c.total);
    executionInfo.insert(name + " [max secs]", c.mx);
    executionInfo.insert(name + " [min secs]", c.mn);
  };
  put("Initialize", m_statInit);
  put("Finalize", m_statFinalize);
  put("One-qubit Gate Total", m_stat1q);
  put("Two-qubit Gate Total", m_stat2q);
  std::vector<double> st(13, 0.0);
  check(mps_stats(m_handle, st.data(), 13), "stats");
  for (size_t i = 0; i < st.size() && i < m_statsAtInit.size(); ++i) st[i] -= m_statsAtInit[i];   // this execute() only
  executionInfo.insert("Contract Two-Qubit Gate Tensor [secs]", st[5] * 1e-3);   // merge GEMM + gate, ExaTnMpsVisitor.cpp:1523
  executionInfo.insert("Decompose Tensor SVD [secs]", st[6] * 1e-3);             // :1626
  executionInfo.insert("Truncate SVD Tensor [secs]", st[7] * 1e-3);              // :1718 (fused with the write-back here)
  executionInfo.insert("Contract Single-Qubit Gate Tensor [calls]", (int)st[1]); // 1q gates that ran as their own kernel (:1256)
  executionInfo.insert("Two-qubit Gate Total [gpu gates]", (int)st[0]);
  executionInfo.insert("b200-gates-2q", st[0]);
  executionInfo.insert("b200-layers", st[2]);
  executionInfo.insert("b200-jacobi-sweeps", st[3]);
  executionInfo.insert("b200-kernel-launches", st[4]);
  executionInfo.insert("b200-svd-nonconverged", st[11]);
  executionInfo.insert("b200-norm-guard-violations", st[12]);
  executionInfo.insert("b200-discarded-weight", discardedWeight());
}

void B200MpsVisitor::finalize() {
  if (!m_handle) xacc::error("B200MpsVisitor::finalize called before initialize");
  StatTimer timer;
  double norm = 0.0;
  check(mps_norm(m_handle, 0, &norm), "norm");
  m_buffer->addExtraInfo("norm", norm);   // reference adds it for n < 20 only (:612); harmless beyond
  if (!m_measureQubits.empty()) {
    if (m_shotCount < 1 && m_nQubits < 20) {
      // "exp-val-z": raw <psi| prod Z |psi>, not divided by the norm (:616-644)
      std::vector<int> q(m_measureQubits.begin(), m_measureQubits.end());
      double ez = 0.0;
      check(mps_expval_z(m_handle, 0, (int)q.size(), q.data(), &ez), "expval_z");
      m_buffer->addExtraInfo("exp-val-z", ez);
    } else if (m_shotCount >= 1) {
      int nm = 0;   // the handle's own measure list sets the stride of a sample string
      check(mps_n_measured(m_handle, &nm), "n_measured");
      std::vector<char> out((size_t)m_shotCount * nm + 1);
      int produced = 0;
      check(mps_sample(m_handle, 0, m_shotCount, out.data(), out.size(), &produced), "sample");
      for (int s = 0; s < produced; ++s) m_buffer->appendMeasurement(std::string(out.data() + (size_t)s * nm, nm));
    }
  }
  // {"bitstring", vector<int>}: amplitude of one bit string, or the normalised wave-function slice over the legs marked -1
  // (ExaTnMpsVisitor.cpp:776-822; the reference offers it in its MPI build only, same buffer keys here)
  if (options.keyExists<std::vector<int>>("bitstring")) {
    const std::vector<int> bitString = options.get<std::vector<int>>("bitstring");
    if ((int)bitString.size() != m_nQubits) xacc::error("Bitstring size must match the number of qubits.");
    int nOpen = 0;
    for (int v : bitString) nOpen += (v < 0);
    if (nOpen > 24) xacc::error("bitstring: more than 24 open legs");
    std::vector<int8_t> b(bitString.begin(), bitString.end());
    std::vector<std::complex<double>> slice((size_t)1 << nOpen);
    size_t len = 0;
    check(mps_amplitude(m_handle, 0, b.data(), reinterpret_cast<double*>(slice.data()), &len), "amplitude");
    if (slice.size() == 1) {
      m_buffer->addExtraInfo("amplitude-real", slice[0].real());
      m_buffer->addExtraInfo("amplitude-imag", slice[0].imag());
    } else {
      double nv = 0.0;
      for (const auto& v : slice) nv += std::norm(v);
      const double scale = nv > 1e-12 ? 1.0 / std::sqrt(nv) : 1.0;   // a slice of zero norm stays as it is (:794-806)
      std::vector<double> re, im;
      re.reserve(slice.size()); im.reserve(slice.size());
      for (const auto& v : slice) { re.push_back(v.real() * scale); im.push_back(v.imag() * scale); }
      m_buffer->addExtraInfo("amplitude-real-vec", re);
      m_buffer->addExtraInfo("amplitude-imag-vec", im);
    }
  }
  m_statFinalize.add(timer.secs());
  exportStats();
}

const double B200MpsVisitor::getExpectationValueZ(std::shared_ptr<CompositeInstruction> function) {
  // VQE mode (TNQVM.cpp:52-92): the ansatz has been applied once; every observable term arrives here as its change-of-basis
  // gates plus Measure instructions.  Like the reference's VQE-capable MPS visitor (ITensorMPSVisitor.cpp:440-464) the ansatz
  // state is cached, the term applied, <Z...Z> taken over the measured qubits, and the ansatz state put back -- here on the
  // device (mps_snapshot / mps_restore), and the expectation by one transfer-matrix sweep instead of the 100000-shot estimate
  // of ExaTnMpsVisitor.cpp:1087-1117.
  if (!m_handle) xacc::error("B200MpsVisitor::getExpectationValueZ called before initialize");
  check(mps_snapshot(m_handle), "snapshot");
  std::vector<int> q;
  InstructionIterator it(function);
  while (it.hasNext()) {
    auto inst = it.next();
    if (!inst->isEnabled()) continue;
    if (inst->name() == "Measure") q.push_back((int)inst->bits()[0]);
    else inst->accept(this);
  }
  double result = 0.0;
  if (!q.empty()) {
    double ez = 0.0, nrm = 1.0;
    check(mps_expval_z(m_handle, 0, (int)q.size(), q.data(), &ez), "expval_z");
    check(mps_norm(m_handle, 0, &nrm), "norm");
    result = nrm > 0 ? ez / nrm : 0.0;   // the sampled estimate of the reference is implicitly normalised
  }
  check(mps_restore(m_handle), "restore");
  return result;
}

const std::vector<std::complex<double>> B200MpsVisitor::getState() {
  if (!m_handle || m_nQubits > 30) return {};
  std::vector<std::complex<double>> sv((size_t)1 << m_nQubits);
  check(mps_statevector(m_handle, 0, reinterpret_cast<double*>(sv.data())), "statevector");
  return sv;
}

std::vector<double> B200MpsVisitor::engineStats() const {
  std::vector<double> st(8, 0.0);
  if (m_handle) check(mps_stats(m_handle, st.data(), 8), "stats");
  return st;
}
std::vector<int> B200MpsVisitor::bondDimensions() const {
  std::vector<int> b(std::max(m_nQubits - 1, 1), 1);
  if (m_handle) check(mps_bond_dims(m_handle, b.data()), "bond_dims");
  b.resize(std::max(m_nQubits - 1, 0));
  return b;
}
double B200MpsVisitor::discardedWeight() const {
  double w = 0.0;
  if (m_handle) check(mps_discarded_weight(m_handle, &w), "discarded_weight");
  return w;
}
std::complex<double> B200MpsVisitor::amplitude(const std::vector<int>& bits) const {
  std::vector<int8_t> b(bits.begin(), bits.end());
  std::complex<double> out(0, 0);
  size_t len = 0;
  for (auto v : b) if (v < 0) xacc::error("amplitude(): open legs are not supported through this accessor");
  check(mps_amplitude(m_handle, 0, b.data(), reinterpret_cast<double*>(&out), &len), "amplitude");
  return out;
}
}  // namespace tnqvm
