// tnqvm::B200MpsVisitor -- the `exatn-mps` visitor service re-implemented over libmps_b200.so (include/mps_b200.h).
//
// Drop-in for tnqvm::ExatnMpsVisitor (tnqvm/visitors/exatn-mps/ExaTnMpsVisitor.hpp:50-139): same service name, same
// option keys ("max-bond-dim" int, "svd-cutoff" double; ExaTnMpsVisitor.cpp:257-271), same visit()/finalize() contract
// and the same AcceleratorBuffer outputs ("norm", "exp-val-z", measurement bit strings; ExaTnMpsVisitor.cpp:576-670).
// All tensor work happens on the GPU behind the C ABI; this class only turns XACC instructions into gate matrices.
#pragma once
#include <array>
#include <complex>
#include <string>
#include <vector>

#include "../../../include/mps_b200.h"
#include "TNQVMVisitorShim.hpp"

namespace tnqvm {

// gate name + parameters -> row-major matrix (tnqvm/base/Gates.hpp:132-334 restated; unknown names give the identity like
// ExatnUtils.cpp:112).  Returns the dimension (2 or 4).
int b200GateMatrix(const std::string& name, const std::vector<double>& params, std::complex<double> out[16]);

class B200MpsVisitor : public TNQVMVisitor {
public:
  B200MpsVisitor();
  ~B200MpsVisitor() override;

  void initialize(std::shared_ptr<AcceleratorBuffer> buffer, int nbShots) override;
  void finalize() override;
  const std::string name() const override { return "exatn-mps"; }
  const std::string description() const override { return "B200-native MPS visitor (libmps_b200, sm_100a)"; }
  std::shared_ptr<TNQVMVisitor> clone() override { return std::make_shared<B200MpsVisitor>(); }

  void visit(Identity&) override {}
  void visit(Hadamard& g) override { applyGate(g); }
  void visit(X& g) override { applyGate(g); }
  void visit(Y& g) override { applyGate(g); }
  void visit(Z& g) override { applyGate(g); }
  void visit(Rx& g) override { applyGate(g); }
  void visit(Ry& g) override { applyGate(g); }
  void visit(Rz& g) override { applyGate(g); }
  void visit(T& g) override { applyGate(g); }
  void visit(Tdg& g) override { applyGate(g); }
  void visit(CPhase& g) override { applyGate(g); }
  void visit(U& g) override { applyGate(g); }
  void visit(CNOT& g) override { applyGate(g); }
  void visit(Swap& g) override;
  void visit(CZ& g) override { applyGate(g); }
  void visit(iSwap& g) override { applyGate(g); }
  void visit(fSim& g) override { applyGate(g); }
  void visit(Measure& g) override;
  // gates XACC's AllGateVisitor would otherwise leave to defaults: handled natively (Gates.hpp has their matrices)
#ifndef TNQVM_B200_WITH_XACC
  void visit(S& g) override { applyGate(g); }
  void visit(Sdg& g) override { applyGate(g); }
  void visit(CY& g) override { applyGate(g); }
  void visit(CH& g) override { applyGate(g); }
  void visit(CRZ& g) override { applyGate(g); }
#endif

  const double getExpectationValueZ(std::shared_ptr<CompositeInstruction> function) override;
  // one ansatz MPS serves all observable terms (TNQVM.cpp:52-92); the reference exatn-mps visitor leaves this false
  bool supportVqeMode() const override { return true; }
  const std::vector<std::complex<double>> getState() override;

  // not part of the reference surface: engine counters for getExecutionInfo()-style reporting
  std::vector<double> engineStats() const;
  std::vector<int> bondDimensions() const;
  double discardedWeight() const;
  std::complex<double> amplitude(const std::vector<int>& bits) const;

private:
  void applyGate(xacc::Instruction& inst);
  void check(int rc, const char* what) const;

  // the reference's per-phase statistics (FunctionCallStat, ExatnUtils.hpp:57-126; buckets named at ExaTnMpsVisitor.cpp:345,
  // 680, 1256, 1292, 1523, 1553, 1626, 1711, 1718, 1729), exported through getExecutionInfo() instead of being printed
  struct CallStat { double total = 0, mx = 0, mn = 0; int calls = 0; void add(double s); };
  CallStat m_statInit, m_statFinalize, m_stat1q, m_stat2q;
  void exportStats();

  mps_handle_t m_handle = nullptr;
  std::vector<int> m_devices;   // placement the handle was created with (the handle is reused across execute() calls when unchanged)
  bool m_byCount = false;
  int m_maxBondAtCreate = 0;
  std::vector<double> m_statsAtInit;   // engine counters when this execute() began (the handle outlives it)
  int m_nQubits = 0;
  int m_shotCount = -1;
  std::vector<size_t> m_measureQubits;
  std::shared_ptr<AcceleratorBuffer> m_buffer;
};
}  // namespace tnqvm
