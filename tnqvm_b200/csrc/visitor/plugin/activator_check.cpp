// Loads the shim-built bundle entry point, starts the activator and checks that exactly one tnqvm::TNQVMVisitor service named
// "exatn-mps" was registered (what TNQVM.hpp:78-92 looks up).  Exit code 0 = ok.  Needs no GPU.
#include <cstdio>
#include <typeindex>
#include "cppmicroservices/BundleActivator.h"
#include "B200MpsVisitor.hpp"
extern "C" cppmicroservices::BundleActivator* b200_shim_create_activator();
extern "C" void b200_shim_destroy_activator(cppmicroservices::BundleActivator*);
int main() {
  cppmicroservices::BundleContext::Registry reg;
  auto* a = b200_shim_create_activator();
  a->Start(cppmicroservices::BundleContext(&reg));
  auto it = reg.find(std::type_index(typeid(tnqvm::TNQVMVisitor)));
  if (it == reg.end() || it->second.size() != 1) { fprintf(stderr, "no TNQVMVisitor service registered\n"); return 1; }
  auto v = std::static_pointer_cast<tnqvm::TNQVMVisitor>(it->second[0]);
  if (v->name() != "exatn-mps") { fprintf(stderr, "unexpected visitor name %s\n", v->name().c_str()); return 1; }
  if (!v->supportVqeMode()) return 1;
  a->Stop(cppmicroservices::BundleContext(&reg));
  b200_shim_destroy_activator(a);
  printf("activator ok: service tnqvm::TNQVMVisitor name=%s\n", v->name().c_str());
  return 0;
}
