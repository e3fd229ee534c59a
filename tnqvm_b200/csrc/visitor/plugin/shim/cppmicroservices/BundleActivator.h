// Minimal stand-in for CppMicroServices (not installed in the development image): just enough of the interface for
// plugin/B200MpsActivator.cpp to be compiled and exercised by the in-tree Makefile and tests.  NOT used in a real XACC build.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeindex>
#include <vector>
#define US_ABI_LOCAL
namespace cppmicroservices {
class BundleContext {
public:
  using Registry = std::map<std::type_index, std::vector<std::shared_ptr<void>>>;
  explicit BundleContext(Registry* r = nullptr) : m_registry(r) {}
  template <class Interface, class Impl>
  void RegisterService(std::shared_ptr<Impl> service) {
    std::shared_ptr<Interface> as_interface = service;   // compile-time check: Impl implements Interface
    if (m_registry) (*m_registry)[std::type_index(typeid(Interface))].push_back(std::static_pointer_cast<void>(as_interface));
  }
private:
  Registry* m_registry;
};
class BundleActivator {
public:
  virtual ~BundleActivator() {}
  virtual void Start(BundleContext context) = 0;
  virtual void Stop(BundleContext context) = 0;
};
}  // namespace cppmicroservices
// the real macro exports create/destroy entry points named after US_BUNDLE_NAME; the shim keeps the same shape
#define CPPMICROSERVICES_EXPORT_BUNDLE_ACTIVATOR(T)                                                         \
  extern "C" cppmicroservices::BundleActivator* b200_shim_create_activator() { return new T(); }            \
  extern "C" void b200_shim_destroy_activator(cppmicroservices::BundleActivator* a) { delete a; }
