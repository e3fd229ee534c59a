"""The reference's own known-answer tests for the exatn-mps path, restated as data
(tnqvm/visitors/exatn-mps/tests/*.cpp).  Each case: qubit count, circuit, and the
measurement-string probabilities the gtest asserts (exact values; the gtests use +-0.05 / +-0.01 on 10 000 shots).
String character i belongs to the i-th Measure (GateMatrixAlgebra.hpp:128-136)."""
import math

H, X, CX, M = "H", "X", "CNOT", "Measure"


def g(name, *qs, p=()):
    return (name, tuple(qs), tuple(p))


def meas(*qs):
    return [g(M, q) for q in qs]


PROB_CASES = []


def case(name, n, circ, expect, cite):
    PROB_CASES.append(dict(name=name, n=n, circuit=circ, expect=expect, cite=cite))


# MpsGateTester.cpp:7-88 checkSimple
for i, s in enumerate(["1000", "0100", "0010", "0001"]):
    case("checkSimple_%d" % i, 4, [g(H, i)] + meas(0, 1, 2, 3), {"0000": 0.5, s: 0.5}, "MpsGateTester.cpp:7-88")
# MpsGateTester.cpp:90-218 checkTwoQubitGates foo1..foo7
case("foo1", 4, [g(H, 0), g(CX, 0, 1)] + meas(0, 1), {"00": 0.5, "11": 0.5}, "MpsGateTester.cpp:92-108")
case("foo2", 4, [g(H, 1), g(CX, 1, 0)] + meas(0, 1), {"00": 0.5, "11": 0.5}, "MpsGateTester.cpp:110-126")
case("foo3", 4, [g(H, 1), g(CX, 1, 2)] + meas(1, 2), {"00": 0.5, "11": 0.5}, "MpsGateTester.cpp:128-144")
case("foo4", 4, [g(H, 2), g(CX, 2, 1)] + meas(2, 1), {"00": 0.5, "11": 0.5}, "MpsGateTester.cpp:146-162")
case("foo5", 4, [g(H, 2), g(CX, 2, 3)] + meas(2, 3), {"00": 0.5, "11": 0.5}, "MpsGateTester.cpp:164-180")
case("foo6", 4, [g(H, 3), g(CX, 3, 2)] + meas(2, 3), {"00": 0.5, "11": 0.5}, "MpsGateTester.cpp:182-198")
case("foo7", 4, [g(X, 0), g(CX, 0, 1), g(CX, 1, 2), g(CX, 2, 3)] + meas(0, 1, 2, 3), {"1111": 1.0}, "MpsGateTester.cpp:200-217")
# MpsGateTester.cpp:220-240 checkDistanceQubitGate (goes through the nearest-neighbour pass)
case("bar1", 4, [g(X, 0), g(CX, 0, 3)] + meas(0, 1, 2, 3), {"1001": 1.0}, "MpsGateTester.cpp:220-240")
# MpsGateTester.cpp:242-310 checkTwoQubits / checkSingleQubit
case("f1", 2, [g(X, 0), g(CX, 0, 1)] + meas(0, 1), {"11": 1.0}, "MpsGateTester.cpp:244-259")
case("f2", 2, [g(X, 1), g(CX, 1, 0)] + meas(0, 1), {"11": 1.0}, "MpsGateTester.cpp:261-276")
case("func1", 1, [g(X, 0)] + meas(0), {"1": 1.0}, "MpsGateTester.cpp:281-294")
case("func2", 1, [g(H, 0), g("Z", 0), g(H, 0)] + meas(0), {"1": 1.0}, "MpsGateTester.cpp:296-309")
# MpsGateTester.cpp:312-331 iSwap
case("iswap1", 3, [g(X, 0), g("iSwap", 0, 1)] + meas(0, 1, 2), {"010": 1.0}, "MpsGateTester.cpp:312-331")
# MpsGateTester.cpp:333-357 fSim: X(q0); fSim(q0,q1,theta,0) -> P("01") = sin^2(theta)
for k in range(10):
    th = -math.pi + k * (2 * math.pi / 9)
    case("fsim_%d" % k, 2, [g(X, 0), g("fSim", 0, 1, p=(th, 0.0))] + meas(0, 1),
         {"01": math.sin(th) ** 2, "10": 1.0 - math.sin(th) ** 2}, "MpsGateTester.cpp:333-357")

# MpsGateTester.cpp:359-407 testDeuteron: <Z0 Z1> vs the 20-entry table
DEUTERON_TABLE = [0.0, -0.324699, -0.614213, -0.837166, -0.9694, -0.996584, -0.915773, -0.735724, -0.475947, -0.164595,
                  0.164595, 0.475947, 0.735724, 0.915773, 0.996584, 0.9694, 0.837166, 0.614213, 0.324699, 0.0]


def deuteron_circuit(t):
    return [g(X, 0), g("Ry", 1, p=(t,)), g(CX, 1, 0), g(H, 0), g(H, 1)] + meas(0, 1)


def deuteron_angles():
    return [-math.pi + k * (2 * math.pi / 19) for k in range(20)]


# MpsGateTester.cpp:409-499 testGrover: P("110") > 0.5, Measure order q2,q1,q0
def grover_circuit():
    T, Td = "T", "Tdg"
    seq = [(H, 0), (H, 1), (H, 2), (X, 0), (H, 2), (H, 2), (CX, 1, 2), (Td, 2), (CX, 0, 2), (T, 2), (CX, 1, 2), (Td, 2), (CX, 0, 2),
           (T, 2), (H, 2), (T, 1), (CX, 0, 1), (T, 0), (Td, 1), (CX, 0, 1), (X, 0), (H, 2), (H, 0), (H, 1), (H, 2), (X, 0), (X, 1),
           (X, 2), (H, 2), (H, 2), (CX, 1, 2), (Td, 2), (CX, 0, 2), (T, 2), (CX, 1, 2), (Td, 2), (CX, 0, 2), (T, 2), (H, 2), (T, 1),
           (CX, 0, 1), (T, 0), (Td, 1), (CX, 0, 1), (H, 2), (X, 0), (X, 1), (X, 2), (H, 0), (H, 1), (H, 2)]
    return [g(s[0], *s[1:]) for s in seq] + meas(2, 1, 0)


# MpsMeasurementTester.cpp:7-35: 35-qubit GHZ, Measure q5,q3,q7,q34 -> only "0000"/"1111" (n >= 20 sampling branch)
def ghz35():
    return [g(H, 0)] + [g(CX, i, i + 1) for i in range(34)] + meas(5, 3, 7, 34)


# MpsMeasurementTester.cpp:37-66 checkRandomSeed: 4-qubit GHZ, seed 123, 8192 shots -> identical counts on every run
def ghz4_measured():
    return [g(H, 0)] + [g(CX, i, i + 1) for i in range(3)] + meas(0, 1, 2, 3)


# ITensorMPSVisitorTester.cpp:322-390 (sibling suite, path independent): <Z> = 1 - 2 sin^2(theta/2) after Rx(theta)
def rx_expz(theta):
    return 1.0 - 2.0 * math.sin(theta / 2.0) ** 2


def probs_from_state(state, n, measured):
    """Exact measurement-string distribution of a dense state (qubit 0 = LSB)."""
    import numpy as np
    out = {}
    p = np.abs(state) ** 2
    for idx in np.nonzero(p > 1e-14)[0]:
        s = "".join("1" if (int(idx) >> q) & 1 else "0" for q in measured)
        out[s] = out.get(s, 0.0) + float(p[idx])
    return out
