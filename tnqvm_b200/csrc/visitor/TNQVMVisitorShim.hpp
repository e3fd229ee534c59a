// The TNQVMVisitor contract the accelerator drives (reference: tnqvm/visitors/TNQVMVisitor.hpp:58-82), declared here
// only when the real TNQVM/XACC headers are not available.  With -DTNQVM_B200_WITH_XACC the reference's own header
// is included instead and this file is empty.
#pragma once
#ifdef TNQVM_B200_WITH_XACC
#include "TNQVMVisitor.hpp"
#else
#include <complex>
#include <memory>
#include <vector>

#include "xacc_shim.hpp"

namespace tnqvm {
using namespace xacc;
using namespace xacc::quantum;

class TNQVMVisitor : public AllGateVisitor, public OptionsProvider, public xacc::Cloneable<TNQVMVisitor> {
public:
  virtual void initialize(std::shared_ptr<AcceleratorBuffer> buffer, int nbShots = 1) = 0;
  virtual const double getExpectationValueZ(std::shared_ptr<CompositeInstruction> function) = 0;
  virtual const std::vector<std::complex<double>> getState() { return {}; }
  virtual void finalize() = 0;
  void setOptions(const HeterogeneousMap& in_options) { options = in_options; }
  virtual void setKernelName(const std::string&) {}
  virtual bool supportVqeMode() const { return false; }
  HeterogeneousMap getExecutionInfo() const { return executionInfo; }

protected:
  std::shared_ptr<AcceleratorBuffer> buffer;
  HeterogeneousMap options;
  HeterogeneousMap executionInfo;
};
}  // namespace tnqvm
#endif
