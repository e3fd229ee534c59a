// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Thin C-ABI wrapper compiled against the REFERENCE'S OWN header-only sources where they lie
// (-I /root/reference/tnqvm): base/Gates.hpp (gate matrices), utils/GateMatrixAlgebra.hpp
// (dense state-vector gate application + GenerateSamples) and utils/RandomEngine.hpp.
// No reference source is copied into this repo; the output goes to oracle/_ref/ (git-ignored).
// Used to validate oracle/mps_oracle.cpp and to generate tests/golden/*.json.
#include <array>
#include <cassert>
#include <complex>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "base/Gates.hpp"
#include "utils/RandomEngine.hpp"
#include "utils/GateMatrixAlgebra.hpp"

using namespace tnqvm;
typedef std::vector<std::vector<std::complex<double>>> Mat;

static Mat ref_matrix(const std::string& name, const double* p) {
  switch (GetGateType(name)) {   // same dispatch as ExatnUtils.cpp:80-113
    case CommonGates::Rx: return GetGateMatrix<CommonGates::Rx>(p[0]);
    case CommonGates::Ry: return GetGateMatrix<CommonGates::Ry>(p[0]);
    case CommonGates::Rz: return GetGateMatrix<CommonGates::Rz>(p[0]);
    case CommonGates::I: return GetGateMatrix<CommonGates::I>();
    case CommonGates::H: return GetGateMatrix<CommonGates::H>();
    case CommonGates::X: return GetGateMatrix<CommonGates::X>();
    case CommonGates::Y: return GetGateMatrix<CommonGates::Y>();
    case CommonGates::Z: return GetGateMatrix<CommonGates::Z>();
    case CommonGates::T: return GetGateMatrix<CommonGates::T>();
    case CommonGates::Tdg: return GetGateMatrix<CommonGates::Tdg>();
    case CommonGates::U: return GetGateMatrix<CommonGates::U>(p[0], p[1], p[2]);
    case CommonGates::CNOT: return GetGateMatrix<CommonGates::CNOT>();
    case CommonGates::CY: return GetGateMatrix<CommonGates::CY>();
    case CommonGates::CZ: return GetGateMatrix<CommonGates::CZ>();
    case CommonGates::CH: return GetGateMatrix<CommonGates::CH>();
    case CommonGates::CRZ: return GetGateMatrix<CommonGates::CRZ>(p[0]);
    case CommonGates::CPhase: return GetGateMatrix<CommonGates::CPhase>(p[0]);
    case CommonGates::Swap: return GetGateMatrix<CommonGates::Swap>();
    case CommonGates::iSwap: return GetGateMatrix<CommonGates::iSwap>();
    case CommonGates::fSim: return GetGateMatrix<CommonGates::fSim>(p[0], p[1]);
    default: return GetGateMatrix<CommonGates::I>();
  }
}

extern "C" int ref_gate_matrix(const char* name, const double* params, double* out_ri) {
  double p[3] = {params ? params[0] : 0, params ? params[1] : 0, params ? params[2] : 0};
  Mat m = ref_matrix(name, p);
  size_t k = 0;
  for (auto& row : m) for (auto& e : row) { out_ri[2 * k] = e.real(); out_ri[2 * k + 1] = e.imag(); ++k; }
  return (int)m.size();
}
extern "C" void ref_apply_1q(double* state_ri, int n, int q, const char* name, const double* params) {
  StateVectorType psi(reinterpret_cast<std::complex<double>*>(state_ri), reinterpret_cast<std::complex<double>*>(state_ri) + (1ULL << n));
  double p[3] = {params ? params[0] : 0, params ? params[1] : 0, params ? params[2] : 0};
  ApplySingleQubitGate(psi, q, ref_matrix(name, p));
  memcpy(state_ri, psi.data(), sizeof(std::complex<double>) * psi.size());
}
extern "C" void ref_apply_cnot(double* state_ri, int n, int ctrl, int tgt) {
  StateVectorType psi(reinterpret_cast<std::complex<double>*>(state_ri), reinterpret_cast<std::complex<double>*>(state_ri) + (1ULL << n));
  ApplyCNOTGate(psi, ctrl, tgt);
  memcpy(state_ri, psi.data(), sizeof(std::complex<double>) * psi.size());
}
extern "C" void ref_set_seed(uint64_t seed) { randomEngine::get_instance().setSeed(seed); }
extern "C" double ref_rand_prob() { return generateRandomProbability(); }
// returns number of strings written (each nbits chars)
extern "C" int ref_generate_samples(const double* state_ri, int n, int shots, const int* bits, int nbits, char* out) {
  std::vector<std::complex<double>> psi(reinterpret_cast<const std::complex<double>*>(state_ri),
                                        reinterpret_cast<const std::complex<double>*>(state_ri) + (1ULL << n));
  std::vector<size_t> mb(bits, bits + nbits);
  auto s = GenerateSamples(psi, (uint64_t)shots, mb);
  for (size_t i = 0; i < s.size(); ++i) memcpy(out + i * nbits, s[i].data(), nbits);
  return (int)s.size();
}
