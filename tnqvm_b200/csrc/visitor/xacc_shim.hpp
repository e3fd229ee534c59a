// Minimal stand-in for the parts of the XACC API the MPS visitor touches:
//   AcceleratorBuffer::size/addExtraInfo/appendMeasurement/getMeasurementCounts,
//   HeterogeneousMap::keyExists<T>/get<T>/insert, Instruction::name/bits/getParameter(i).as<double>/accept,
//   CompositeInstruction / InstructionIterator, the gate classes of xacc::quantum and AllGateVisitor.
// XACC itself (eclipse/xacc, unpinned: the reference CI clones HEAD, .github/workflows/build.yml:30) is not
// installed in this image, so the adapter is compiled and tested against this shim; building with
// -DTNQVM_B200_WITH_XACC switches every include to the real headers.
// Nothing here is product logic: it only gives the visitor the types the real framework would hand it.
#pragma once
#ifdef TNQVM_B200_WITH_XACC
#include "AllGateVisitor.hpp"
#include "Identifiable.hpp"
#include "xacc.hpp"
#else
#include <any>
#include <complex>
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace xacc {

inline bool verbose = false;
inline void info(const std::string& m) { if (verbose) std::cerr << "[xacc info] " << m << "\n"; }
inline void warning(const std::string& m) { std::cerr << "[xacc warning] " << m << "\n"; }
[[noreturn]] inline void error(const std::string& m) { throw std::runtime_error(m); }

class HeterogeneousMap {
public:
  HeterogeneousMap() = default;
  template <typename T> void insert(const std::string& k, const T& v) { items[k] = v; }
  template <typename T> bool keyExists(const std::string& k) const {
    auto it = items.find(k);
    return it != items.end() && std::any_cast<T>(&it->second) != nullptr;
  }
  template <typename T> T get(const std::string& k) const {
    auto it = items.find(k);
    if (it == items.end()) error("HeterogeneousMap: missing key " + k);
    return std::any_cast<T>(it->second);
  }
  bool stringExists(const std::string& k) const { return keyExists<std::string>(k); }
  std::string getString(const std::string& k) const { return get<std::string>(k); }
  void merge(const HeterogeneousMap& o) { for (auto& kv : o.items) items[kv.first] = kv.second; }
  const std::map<std::string, std::any>& raw() const { return items; }   // shim only: lets the driver print what a visitor exported
private:
  std::map<std::string, std::any> items;
};

using ExtraInfo = std::variant<int, double, std::string, std::vector<int>, std::vector<double>, std::vector<std::string>>;

class AcceleratorBuffer {
public:
  AcceleratorBuffer(const std::string& name, int n) : bufferId(name), nBits(n) {}
  int size() const { return nBits; }
  const std::string& name() const { return bufferId; }
  void addExtraInfo(const std::string& key, ExtraInfo v) { info[key] = std::move(v); }
  bool hasExtraInfoKey(const std::string& key) const { return info.count(key) != 0; }
  const ExtraInfo& operator[](const std::string& key) const { return info.at(key); }
  const std::map<std::string, ExtraInfo>& getInformation() const { return info; }
  void appendMeasurement(const std::string& bits) { counts[bits] += 1; }
  void appendMeasurement(const std::string& bits, int c) { counts[bits] += c; }
  const std::map<std::string, int>& getMeasurementCounts() const { return counts; }
  double computeMeasurementProbability(const std::string& bits) const {
    long tot = 0;
    for (auto& kv : counts) tot += kv.second;
    auto it = counts.find(bits);
    return (tot == 0 || it == counts.end()) ? 0.0 : double(it->second) / double(tot);
  }
  void resetBuffer() { info.clear(); counts.clear(); }
private:
  std::string bufferId;
  int nBits;
  std::map<std::string, ExtraInfo> info;
  std::map<std::string, int> counts;
};

class InstructionParameter {
public:
  InstructionParameter(double v = 0.0) : val(v) {}
  template <typename T> T as() const { return static_cast<T>(val); }
private:
  double val;
};

class BaseInstructionVisitor {
public:
  virtual ~BaseInstructionVisitor() = default;
};
template <typename T> class InstructionVisitor {
public:
  virtual void visit(T&) = 0;
  virtual ~InstructionVisitor() = default;
};

class Instruction {
public:
  Instruction(std::string n, std::vector<std::size_t> b, std::vector<InstructionParameter> p = {})
      : gateName(std::move(n)), qbits(std::move(b)), params(std::move(p)) {}
  virtual ~Instruction() = default;
  const std::string name() const { return gateName; }
  const std::vector<std::size_t> bits() const { return qbits; }
  void setBits(const std::vector<std::size_t>& b) { qbits = b; }
  InstructionParameter getParameter(std::size_t i) const { return params.at(i); }
  std::vector<InstructionParameter> getParameters() const { return params; }
  int nParameters() const { return (int)params.size(); }
  int nRequiredBits() const { return (int)qbits.size(); }
  bool isEnabled() const { return enabled; }
  void disable() { enabled = false; }
  virtual bool isComposite() const { return false; }
  virtual void accept(BaseInstructionVisitor* v) = 0;
  void accept(std::shared_ptr<BaseInstructionVisitor> v) { accept(v.get()); }
  std::string toString() const {
    std::string s = gateName + "(";
    for (std::size_t i = 0; i < qbits.size(); ++i) s += (i ? ",q" : "q") + std::to_string(qbits[i]);
    return s + ")";
  }
protected:
  std::string gateName;
  std::vector<std::size_t> qbits;
  std::vector<InstructionParameter> params;
  bool enabled = true;
};

namespace quantum {
// one class per gate, dispatching to InstructionVisitor<Gate>::visit like XACC's DEFINE_VISITABLE()
#define B200_SHIM_GATE(CLS, NAME)                                                                              \
  class CLS : public Instruction {                                                                             \
  public:                                                                                                      \
    CLS(std::vector<std::size_t> b, std::vector<InstructionParameter> p = {}) : Instruction(NAME, std::move(b), std::move(p)) {} \
    using Instruction::accept;                                                                                 \
    void accept(BaseInstructionVisitor* v) override {                                                          \
      auto* c = dynamic_cast<InstructionVisitor<CLS>*>(v);                                                     \
      if (c) c->visit(*this);                                                                                  \
      else xacc::warning(std::string("visitor does not handle ") + NAME);                                      \
    }                                                                                                          \
  };
B200_SHIM_GATE(Identity, "I")
B200_SHIM_GATE(Hadamard, "H")
B200_SHIM_GATE(X, "X")
B200_SHIM_GATE(Y, "Y")
B200_SHIM_GATE(Z, "Z")
B200_SHIM_GATE(Rx, "Rx")
B200_SHIM_GATE(Ry, "Ry")
B200_SHIM_GATE(Rz, "Rz")
B200_SHIM_GATE(T, "T")
B200_SHIM_GATE(Tdg, "Tdg")
B200_SHIM_GATE(S, "S")
B200_SHIM_GATE(Sdg, "Sdg")
B200_SHIM_GATE(U, "U")
B200_SHIM_GATE(CPhase, "CPhase")
B200_SHIM_GATE(CNOT, "CNOT")
B200_SHIM_GATE(Swap, "Swap")
B200_SHIM_GATE(CZ, "CZ")
B200_SHIM_GATE(CY, "CY")
B200_SHIM_GATE(CH, "CH")
B200_SHIM_GATE(CRZ, "CRZ")
B200_SHIM_GATE(iSwap, "iSwap")
B200_SHIM_GATE(fSim, "fSim")
B200_SHIM_GATE(Measure, "Measure")
#undef B200_SHIM_GATE

class AllGateVisitor : public BaseInstructionVisitor,
                       public InstructionVisitor<Identity>, public InstructionVisitor<Hadamard>, public InstructionVisitor<X>,
                       public InstructionVisitor<Y>, public InstructionVisitor<Z>, public InstructionVisitor<Rx>,
                       public InstructionVisitor<Ry>, public InstructionVisitor<Rz>, public InstructionVisitor<T>,
                       public InstructionVisitor<Tdg>, public InstructionVisitor<S>, public InstructionVisitor<Sdg>,
                       public InstructionVisitor<U>, public InstructionVisitor<CPhase>, public InstructionVisitor<CNOT>,
                       public InstructionVisitor<Swap>, public InstructionVisitor<CZ>, public InstructionVisitor<CY>,
                       public InstructionVisitor<CH>, public InstructionVisitor<CRZ>, public InstructionVisitor<iSwap>,
                       public InstructionVisitor<fSim>, public InstructionVisitor<Measure> {
public:
  // gates the reference visitor does not override fall through to defaults in XACC; here: warn once per call
  void visit(S& g) override { unsupported(g); }
  void visit(Sdg& g) override { unsupported(g); }
  void visit(CY& g) override { unsupported(g); }
  void visit(CH& g) override { unsupported(g); }
  void visit(CRZ& g) override { unsupported(g); }
protected:
  virtual void unsupported(Instruction& g) { xacc::warning("gate " + g.name() + " is not handled by this visitor"); }
};
}  // namespace quantum

class CompositeInstruction {
public:
  explicit CompositeInstruction(std::string n) : kernelName(std::move(n)) {}
  const std::string name() const { return kernelName; }
  void addInstruction(std::shared_ptr<Instruction> i) { insts.push_back(std::move(i)); }
  void addInstructions(const std::vector<std::shared_ptr<Instruction>>& v) { insts.insert(insts.end(), v.begin(), v.end()); }
  int nInstructions() const { return (int)insts.size(); }
  std::shared_ptr<Instruction> getInstruction(int i) const { return insts.at(i); }
  const std::vector<std::shared_ptr<Instruction>>& getInstructions() const { return insts; }
  void clear() { insts.clear(); }
private:
  std::string kernelName;
  std::vector<std::shared_ptr<Instruction>> insts;
};

class InstructionIterator {
public:
  explicit InstructionIterator(std::shared_ptr<CompositeInstruction> k) : kernel(std::move(k)) {}
  bool hasNext() const { return pos < kernel->nInstructions(); }
  std::shared_ptr<Instruction> next() { return kernel->getInstruction(pos++); }
private:
  std::shared_ptr<CompositeInstruction> kernel;
  int pos = 0;
};

class OptionsProvider {
public:
  virtual ~OptionsProvider() = default;
};
class Identifiable {
public:
  virtual const std::string name() const = 0;
  virtual const std::string description() const = 0;
  virtual ~Identifiable() = default;
};
template <typename T> class Cloneable : public Identifiable {
public:
  virtual std::shared_ptr<T> clone() = 0;
};

// gate factory (the subset of IRProvider::createInstruction the nearest-neighbour pass and the XASM reader need)
std::shared_ptr<Instruction> createInstruction(const std::string& name, const std::vector<std::size_t>& bits,
                                               const std::vector<InstructionParameter>& params = {});
}  // namespace xacc
#endif
