// CppMicroServices bundle activator of the B200 MPS visitor: makes `xacc::getAccelerator("tnqvm", {{"tnqvm-visitor", "exatn-mps"}})`
// find tnqvm::B200MpsVisitor (name() == "exatn-mps", TNQVM.hpp:78-92) the way it finds the reference visitor
// (tnqvm/visitors/exatn-mps/ExaTnMpsActivator.cpp:14-19 registers its visitor as a tnqvm::TNQVMVisitor service).
// Built by plugin/CMakeLists.txt inside an XACC/TNQVM tree (-DTNQVM_B200_WITH_XACC); the in-tree Makefile only syntax-checks it
// against plugin/shim/ (XACC and CppMicroServices are not installed in the development image).
#include "cppmicroservices/BundleActivator.h"
#include "cppmicroservices/BundleContext.h"
#include "cppmicroservices/ServiceProperties.h"

#include "B200MpsVisitor.hpp"

using namespace cppmicroservices;

class US_ABI_LOCAL B200MpsActivator : public BundleActivator {
public:
  B200MpsActivator() {}

  void Start(BundleContext context) override {
    // only the visitor: the nearest-neighbour IRTransformation and the RCS generator of the reference bundle are XACC-side
    // services that stay where they are (out of scope)
    context.RegisterService<tnqvm::TNQVMVisitor>(std::make_shared<tnqvm::B200MpsVisitor>());
  }

  void Stop(BundleContext /*context*/) override {}
};

CPPMICROSERVICES_EXPORT_BUNDLE_ACTIVATOR(B200MpsActivator)
