// b200_tnqvm_run -- drives tnqvm::B200MpsVisitor exactly the way the accelerator does (reference: TNQVM::execute,
// tnqvm/TNQVM.cpp:106-139): setOptions -> initialize(buffer, shots) -> nearest-neighbour rewrite (max-distance 1)
// -> accept(visitor) for every enabled instruction -> finalize -> results in the AcceleratorBuffer.
// It stands in for the XACC runtime, which is not installed here: circuits come from the XASM subset the reference's
// tests and examples/sycamore/resources/*.xasm use.  Output: one JSON object on stdout.
//
// The nearest-neighbour pass restates the meet-in-the-middle Swap ladder of "lnn-transform"
// (tnqvm/visitors/exatn-mps/NearestNeighborTransform.hpp:43-135); the accelerator itself calls XACC's external "nnizer".
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "B200MpsVisitor.hpp"

namespace xacc {
std::shared_ptr<Instruction> createInstruction(const std::string& nm, const std::vector<std::size_t>& bits,
                                               const std::vector<InstructionParameter>& params) {
  using namespace xacc::quantum;
#define MK(N, CLS) if (nm == N) return std::make_shared<CLS>(bits, params);
  MK("I", Identity) MK("H", Hadamard) MK("X", X) MK("Y", Y) MK("Z", Z) MK("Rx", Rx) MK("Ry", Ry) MK("Rz", Rz) MK("T", T)
  MK("Tdg", Tdg) MK("S", S) MK("Sdg", Sdg) MK("U", U) MK("U3", U) MK("CPhase", CPhase) MK("CNOT", CNOT) MK("CX", CNOT)
  MK("Swap", Swap) MK("CZ", CZ) MK("CY", CY) MK("CH", CH) MK("CRZ", CRZ) MK("iSwap", iSwap) MK("fSim", fSim)
  MK("Measure", Measure)
#undef MK
  xacc::error("unknown instruction: " + nm);
}
}  // namespace xacc

namespace {
using namespace xacc;

// tiny arithmetic evaluator for XASM parameters: numbers, pi, + - * /, parentheses, unary minus
struct Expr {
  const char* p;
  double num() {
    while (isspace(*p)) ++p;
    if (*p == '(') { ++p; double v = sum(); while (isspace(*p)) ++p; if (*p == ')') ++p; return v; }
    if (*p == '-') { ++p; return -num(); }
    if (*p == '+') { ++p; return num(); }
    if (!strncmp(p, "pi", 2)) { p += 2; return M_PI; }
    char* e; double v = strtod(p, &e);
    if (e == p) xacc::error(std::string("cannot parse parameter near: ") + p);
    p = e; return v;
  }
  double prod() { double v = num(); for (;;) { while (isspace(*p)) ++p; if (*p == '*') { ++p; v *= num(); } else if (*p == '/') { ++p; v /= num(); } else return v; } }
  double sum() { double v = prod(); for (;;) { while (isspace(*p)) ++p; if (*p == '+') { ++p; v += prod(); } else if (*p == '-') { ++p; v -= prod(); } else return v; } }
};

std::shared_ptr<CompositeInstruction> parseXasm(const std::string& text, int& nQubitsSeen) {
  auto kernel = std::make_shared<CompositeInstruction>("kernel");
  std::istringstream in(text);
  std::string line;
  nQubitsSeen = 0;
  while (std::getline(in, line)) {
    const auto c = line.find("//");
    if (c != std::string::npos) line = line.substr(0, c);
    const auto lp = line.find('('), rp = line.rfind(')');
    if (lp == std::string::npos || rp == std::string::npos || line.find("__qpu__") != std::string::npos) continue;
    std::string name = line.substr(0, lp);
    name.erase(0, name.find_first_not_of(" \t"));
    name.erase(name.find_last_not_of(" \t") + 1);
    if (name.empty()) continue;
    std::vector<std::size_t> bits;
    std::vector<InstructionParameter> params;
    std::string args = line.substr(lp + 1, rp - lp - 1), a;
    int depth = 0;
    std::vector<std::string> parts;
    for (char ch : args) {
      if (ch == '(') ++depth;
      if (ch == ')') --depth;
      if (ch == ',' && depth == 0) { parts.push_back(a); a.clear(); } else a += ch;
    }
    if (!a.empty()) parts.push_back(a);
    for (auto& s : parts) {
      const auto lb = s.find('['), rb = s.find(']');
      const auto first = s.find_first_not_of(" \t");
      if (lb != std::string::npos && rb != std::string::npos && first != std::string::npos && (isalpha(s[first]) || s[first] == '_') &&
          s.substr(first, 2) != "pi") {
        bits.push_back((std::size_t)atoi(s.substr(lb + 1, rb - lb - 1).c_str()));
      } else if (first != std::string::npos) {
        Expr e{s.c_str()};
        params.emplace_back(e.sum());
      }
    }
    for (auto b : bits) nQubitsSeen = std::max(nQubitsSeen, (int)b + 1);
    kernel->addInstruction(createInstruction(name, bits, params));
  }
  return kernel;
}

void nearestNeighborTransform(std::shared_ptr<CompositeInstruction> program, int maxDistance = 1) {
  std::vector<std::shared_ptr<Instruction>> out;
  auto far = [&](long a, long b) { return std::labs(a - b) > maxDistance; };
  for (auto& inst : program->getInstructions()) {
    const auto bits = inst->bits();
    if (bits.size() == 2 && far((long)bits[0], (long)bits[1])) {
      const std::size_t lo0 = std::min(bits[0], bits[1]), hi0 = std::max(bits[0], bits[1]);
      std::size_t lo = lo0, hi = hi0;
      for (;;) {
        out.push_back(createInstruction("Swap", {lo, lo + 1}));
        ++lo;
        if (!far((long)lo, (long)hi)) break;
        out.push_back(createInstruction("Swap", {hi, hi - 1}));
        --hi;
        if (!far((long)lo, (long)hi)) break;
      }
      inst->setBits(bits[0] < bits[1] ? std::vector<std::size_t>{lo, hi} : std::vector<std::size_t>{hi, lo});
      out.push_back(inst);
      for (std::size_t i = lo; i > lo0; --i) out.push_back(createInstruction("Swap", {i, i - 1}));
      for (std::size_t i = hi; i < hi0; ++i) out.push_back(createInstruction("Swap", {i, i + 1}));
    } else {
      out.push_back(inst);
    }
  }
  program->clear();
  program->addInstructions(out);
}

// TNQVM::execute, tnqvm/TNQVM.cpp:106-139
void execute(std::shared_ptr<tnqvm::TNQVMVisitor> visitor, const HeterogeneousMap& options, std::shared_ptr<AcceleratorBuffer> buffer,
             std::shared_ptr<CompositeInstruction> kernel, int shots) {
  visitor->setOptions(options);
  visitor->initialize(buffer, shots);
  visitor->setKernelName(kernel->name());
  if (visitor->name() == "exatn-mps") nearestNeighborTransform(kernel, 1);
  InstructionIterator it(kernel);
  while (it.hasNext()) {
    auto inst = it.next();
    if (inst->isEnabled()) inst->accept(visitor);
  }
  visitor->finalize();
}

// TNQVM::execute(buffer, vector<kernels>) in VQE mode, tnqvm/TNQVM.cpp:52-92: the base (ansatz) kernel once, then one
// getExpectationValueZ per observed sub-circuit.  A term is a Pauli word such as "X0X1" or "Z3"; its sub-circuit is what
// XACC's observe() appends: H for X, Rx(pi/2) for Y, nothing for Z, then Measure on every qubit of the word.
std::vector<double> executeVqe(std::shared_ptr<tnqvm::TNQVMVisitor> visitor, const HeterogeneousMap& options, std::shared_ptr<AcceleratorBuffer> buffer,
                               std::shared_ptr<CompositeInstruction> ansatz, const std::vector<std::string>& terms, int shots) {
  if (!visitor->supportVqeMode()) xacc::error("visitor does not support VQE mode");
  visitor->setOptions(options);
  if (visitor->name() == "exatn-mps") nearestNeighborTransform(ansatz, 1);
  visitor->initialize(buffer, shots);
  visitor->setKernelName(ansatz->name());
  InstructionIterator it(ansatz);
  while (it.hasNext()) {
    auto inst = it.next();
    if (inst->isEnabled()) inst->accept(visitor);
  }
  std::vector<double> out;
  for (const auto& term : terms) {
    auto obs = std::make_shared<CompositeInstruction>(term);
    std::vector<std::size_t> measured;
    for (size_t i = 0; i < term.size();) {
      const char pauli = term[i++];
      size_t j = i;
      while (j < term.size() && isdigit((unsigned char)term[j])) ++j;
      if (j == i || (pauli != 'X' && pauli != 'Y' && pauli != 'Z')) xacc::error("cannot parse observable term: " + term);
      const std::size_t q = (std::size_t)atoi(term.substr(i, j - i).c_str());
      i = j;
      if (pauli == 'X') obs->addInstruction(createInstruction("H", {q}));
      if (pauli == 'Y') obs->addInstruction(createInstruction("Rx", {q}, {InstructionParameter(M_PI / 2)}));
      measured.push_back(q);
    }
    for (auto q : measured) obs->addInstruction(createInstruction("Measure", {q}));
    out.push_back(visitor->getExpectationValueZ(obs));
  }
  visitor->finalize();
  return out;
}
}  // namespace

int main(int argc, char** argv) {
  std::string file;
  int nq = 0, shots = -1, maxBond = 0, seed = -1, device = 0, gauge = 0;
  double cutoff = -1.0;
  bool wantState = false, dumpNN = false, fuse2q = false, profile = false;
  int repeat = 1;   // run execute() this many times on the same visitor instance (TNQVM.cpp:109 reuses it); every wall time is reported
  std::vector<double> executeMs;
  std::string bitstring, observe, devices;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return std::string(argv[++i]); };
    if (a == "--xasm") file = next();
    else if (a == "--qubits") nq = atoi(next().c_str());
    else if (a == "--shots") shots = atoi(next().c_str());
    else if (a == "--max-bond-dim") maxBond = atoi(next().c_str());
    else if (a == "--svd-cutoff") cutoff = atof(next().c_str());
    else if (a == "--seed") seed = atoi(next().c_str());
    else if (a == "--device") device = atoi(next().c_str());
    else if (a == "--gauge") gauge = atoi(next().c_str());
    else if (a == "--state") wantState = true;
    else if (a == "--dump-nn") dumpNN = true;   // print the nearest-neighbourised program and exit (no GPU needed)
    else if (a == "--bitstring") bitstring = next();
    else if (a == "--fuse-2q") fuse2q = true;
    else if (a == "--devices") devices = next();   // comma-separated GPU indices: site-sharded run
    else if (a == "--profile") profile = true;   // per-phase GPU timings in the execution info (getExecutionInfo())
    else if (a == "--repeat") repeat = std::max(1, atoi(next().c_str()));
    else if (a == "--observe") observe = next();   // VQE mode: semicolon-separated Pauli words, e.g. "X0X1;Y0Y1;Z0;Z1"
    else { fprintf(stderr, "usage: b200_tnqvm_run --xasm FILE|- [--qubits N] [--shots S] [--max-bond-dim D] [--svd-cutoff E] [--seed K] [--state] [--bitstring 01x1..] [--fuse-2q] [--observe \"X0X1;Z0\"]\n"); return 2; }
  }
  try {
    std::stringstream ss;
    if (file == "-" || file.empty()) ss << std::cin.rdbuf();
    else { std::ifstream f(file); if (!f) xacc::error("cannot open " + file); ss << f.rdbuf(); }
    int seen = 0;
    auto kernel = parseXasm(ss.str(), seen);
    if (nq <= 0) nq = seen;
    if (dumpNN) {
      nearestNeighborTransform(kernel, 1);
      for (auto& inst : kernel->getInstructions()) {
        printf("%s", inst->name().c_str());
        for (auto b : inst->bits()) printf(" %zu", b);
        for (auto& p : inst->getParameters()) printf(" %.17g", p.as<double>());
        printf("\n");
      }
      return 0;
    }
    if (shots >= 0 && shots < 1) xacc::error("Invalid 'shots' parameter.");   // TNQVM.hpp:104-107
    HeterogeneousMap opts;
    opts.insert("tnqvm-visitor", std::string("exatn-mps"));
    if (maxBond > 0) opts.insert("max-bond-dim", maxBond);
    if (cutoff >= 0) opts.insert("svd-cutoff", cutoff);
    if (seed >= 0) opts.insert("seed", seed);
    if (!bitstring.empty()) {   // '0'/'1' fix a leg, any other character ('x', '-') leaves it open
      std::vector<int> bits;
      for (char c : bitstring) bits.push_back(c == '0' ? 0 : c == '1' ? 1 : -1);
      opts.insert("bitstring", bits);
    }
    if (fuse2q) opts.insert("b200-fuse-2q", true);
    if (profile) opts.insert("b200-profile", true);
    if (!devices.empty()) {
      std::vector<int> dv;
      std::stringstream ds(devices);
      for (std::string t; std::getline(ds, t, ',');) if (!t.empty()) dv.push_back(atoi(t.c_str()));
      opts.insert("b200-devices", dv);
    }
    opts.insert("b200-device", device);
    opts.insert("b200-gauge", gauge);
    auto visitor = std::make_shared<tnqvm::B200MpsVisitor>();
    auto buffer = std::make_shared<AcceleratorBuffer>("q", nq);
    const int nInst = kernel->nInstructions();
    std::vector<double> vqeTerms;
    if (!observe.empty()) {
      std::vector<std::string> terms;
      std::stringstream ts(observe);
      for (std::string t; std::getline(ts, t, ';');) if (!t.empty()) terms.push_back(t);
      vqeTerms = executeVqe(visitor, opts, buffer, kernel, terms, shots);
    } else {
      for (int r = 0; r < repeat; ++r) {
        if (r) buffer = std::make_shared<AcceleratorBuffer>("q", nq);
        const auto t0 = std::chrono::steady_clock::now();
        execute(visitor, opts, buffer, kernel, shots);
        executeMs.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
    }
    printf("{\"visitor\": \"%s\", \"qubits\": %d, \"instructions\": %d, \"instructions_after_nn\": %d", visitor->name().c_str(), nq, nInst,
           kernel->nInstructions());
    for (auto& kv : buffer->getInformation())
      if (std::holds_alternative<double>(kv.second)) printf(", \"%s\": %.17g", kv.first.c_str(), std::get<double>(kv.second));
    printf(", \"counts\": {");
    bool first = true;
    for (auto& kv : buffer->getMeasurementCounts()) { printf("%s\"%s\": %d", first ? "" : ", ", kv.first.c_str(), kv.second); first = false; }
    printf("}, \"bond_dims\": [");
    auto bd = visitor->bondDimensions();
    for (size_t i = 0; i < bd.size(); ++i) printf("%s%d", i ? ", " : "", bd[i]);
    printf("], \"discarded_weight\": %.17g", visitor->discardedWeight());
    if (buffer->hasExtraInfoKey("amplitude-real"))   // {"bitstring", ...} with every leg fixed
      printf(", \"amplitude\": [%.17g, %.17g]", std::get<double>((*buffer)["amplitude-real"]), std::get<double>((*buffer)["amplitude-imag"]));
    if (buffer->hasExtraInfoKey("amplitude-real-vec")) {
      const auto& re = std::get<std::vector<double>>((*buffer)["amplitude-real-vec"]);
      const auto& im = std::get<std::vector<double>>((*buffer)["amplitude-imag-vec"]);
      printf(", \"amplitude_slice\": [");
      for (size_t i = 0; i < re.size(); ++i) printf("%s[%.17g, %.17g]", i ? ", " : "", re[i], im[i]);
      printf("]");
    }
    if (!executeMs.empty()) {
      printf(", \"execute_ms\": [");
      for (size_t i = 0; i < executeMs.size(); ++i) printf("%s%.3f", i ? ", " : "", executeMs[i]);
      printf("]");
    }
    if (!observe.empty()) {
      printf(", \"vqe_terms\": [");
      for (size_t i = 0; i < vqeTerms.size(); ++i) printf("%s%.17g", i ? ", " : "", vqeTerms[i]);
      printf("]");
    }
    if (wantState) {
      auto sv = visitor->getState();
      printf(", \"state\": [");
      for (size_t i = 0; i < sv.size(); ++i) printf("%s[%.17g, %.17g]", i ? ", " : "", sv[i].real(), sv[i].imag());
      printf("]");
    }
    {   // TNQVMVisitor::getExecutionInfo(): the reference's stat bucket names + engine counters
      printf(", \"execution_info\": {");
      bool firstInfo = true;
      const HeterogeneousMap execInfo = visitor->getExecutionInfo();   // returned by value
      for (auto& kv : execInfo.raw()) {
        if (auto d = std::any_cast<double>(&kv.second)) { printf("%s\"%s\": %.9g", firstInfo ? "" : ", ", kv.first.c_str(), *d); firstInfo = false; }
        else if (auto i = std::any_cast<int>(&kv.second)) { printf("%s\"%s\": %d", firstInfo ? "" : ", ", kv.first.c_str(), *i); firstInfo = false; }
      }
      printf("}");
    }
    auto st = visitor->engineStats();
    printf(", \"stats\": {\"gates_2q\": %.0f, \"layers\": %.0f, \"jacobi_sweeps\": %.0f, \"launches\": %.0f}}\n", st[0], st[2], st[3], st[4]);
  } catch (const std::exception& e) {
    fprintf(stderr, "b200_tnqvm_run: %s\n", e.what());
    return 1;
  }
  return 0;
}
