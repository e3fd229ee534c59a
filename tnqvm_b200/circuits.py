"""Seeded synthetic circuits for the five BASELINE.json configs, an XASM-subset reader/writer and
the nearest-neighbour rewrite.

A circuit is a list of gates ``(name, qubits, params)`` with XACC gate names
(H X Y Z T Tdg Rx Ry Rz U CNOT/CX CZ CPhase Swap iSwap fSim Measure).  ``to_xasm`` emits text that
the real TNQVM/XACC ``xasm`` compiler accepts, so the same circuit can be run on the reference
elsewhere.

Reference shapes restated here (nothing is copied):
  * rcs layer structure / 1q gate set: tnqvm/visitors/exatn-mps/RandomCircuitGen.hpp:64-110
  * nearest-neighbour rewrite ("lnn-transform", stand-in for XACC's external "nnizer" called at
    tnqvm/TNQVM.cpp:119-124): tnqvm/visitors/exatn-mps/NearestNeighborTransform.hpp:43-135
  * Sycamore XASM input: examples/sycamore/resources/*.xasm (Rx/Ry/Rz/fSim)
"""
import ast
import math
import operator
import re

import numpy as np

RCS_GATES = ["H", "X", "Y", "Z", "T", "Rx", "Ry", "Rz"]   # RandomCircuitGen.hpp:64-66


def ghz(n):
    return [("H", (0,), ())] + [("CNOT", (i, i + 1), ()) for i in range(n - 1)]


def brickwork(n, depth, seed=12345, two_qubit="CNOT", prefix_ghz=False):
    """C1/C2: per layer one random 1q gate per qubit from the rcs set (angle U(-pi,pi)), then the
    2q gate on pairs (2j,2j+1) for even layers / (2j+1,2j+2) for odd layers."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = ghz(n) if prefix_ghz else []
    for layer in range(depth):
        for q in range(n):
            g = RCS_GATES[int(rng.integers(0, len(RCS_GATES)))]
            if g in ("Rx", "Ry", "Rz"):
                c.append((g, (q,), (float(rng.uniform(-math.pi, math.pi)),)))
            else:
                c.append((g, (q,), ()))
        start = layer % 2
        for j in range(start, n - 1, 2):
            c.append((two_qubit, (j, j + 1), ()))
    return c


def rcs(n, nlayers, seed=0):
    """Shape of the reference's `rcs` circuit (RandomCircuitGen.hpp:96-118): random 1q layer, then a
    full CX ladder j -> j+1, repeated; Measure on every qubit at the end.  Seeded (the reference
    uses unseeded std::rand)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = []
    for _ in range(nlayers):
        for q in range(n):
            g = RCS_GATES[int(rng.integers(0, len(RCS_GATES)))]
            if g in ("Rx", "Ry", "Rz"):
                c.append((g, (q,), (float(rng.uniform(-math.pi, math.pi)),)))
            else:
                c.append((g, (q,), ()))
        for j in range(n - 1):
            c.append(("CNOT", (j, j + 1), ()))
    for q in range(n):
        c.append(("Measure", (q,), ()))
    return c


def qaoa_ring(n, p, seed=7):
    """C3: ring MaxCut QAOA: H on all; per layer ZZ(gamma) = CX.Rz(2 gamma).CX on the n ring edges
    (the wrap edge (n-1,0) is long-range: nearest_neighbor() routes it), then Rx(2 beta) on all."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gammas = rng.uniform(0, math.pi, p)
    betas = rng.uniform(0, math.pi / 2, p)
    c = [("H", (q,), ()) for q in range(n)]
    for l in range(p):
        edges = [(i, i + 1) for i in range(0, n - 1, 2)] + [(i, i + 1) for i in range(1, n - 1, 2)] + [(n - 1, 0)]
        for (a, b) in edges:
            c.append(("CNOT", (a, b), ()))
            c.append(("Rz", (b,), (float(2 * gammas[l]),)))
            c.append(("CNOT", (a, b), ()))
        for q in range(n):
            c.append(("Rx", (q,), (float(2 * betas[l]),)))
    return c


def hea(n, layers, seed=0):
    """C4: hardware-efficient ansatz: per layer Ry, Rz on every qubit then a CX ladder."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = []
    for _ in range(layers):
        for q in range(n):
            c.append(("Ry", (q,), (float(rng.uniform(-math.pi, math.pi)),)))
            c.append(("Rz", (q,), (float(rng.uniform(-math.pi, math.pi)),)))
        for j in range(n - 1):
            c.append(("CNOT", (j, j + 1), ()))
    return c


def sycamore_like(n, depth, seed=3):
    """C5 stand-in when /root/reference is absent (GPU box): Sycamore-style layers of
    sqrt-X/sqrt-Y/sqrt-W-like 1q rotations followed by fSim(pi/2, pi/6) on a 1D pairing."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = []
    for layer in range(depth):
        for q in range(n):
            k = int(rng.integers(0, 3))
            if k == 0:
                c.append(("Rx", (q,), (math.pi / 2,)))
            elif k == 1:
                c.append(("Ry", (q,), (math.pi / 2,)))
            else:
                c += [("Rz", (q,), (-math.pi / 4,)), ("Rx", (q,), (math.pi / 2,)), ("Rz", (q,), (math.pi / 4,))]
        for j in range(layer % 2, n - 1, 2):
            c.append(("fSim", (j, j + 1), (math.pi / 2, math.pi / 6)))
    return c


def sycamore_grid(depth=14, rows=9, cols=6, n=53, seed=0):
    """C5 at full shape without the reference's resource file (absent on the GPU box): a 53-qubit 2D random circuit in the
    style of examples/sycamore/resources/sycamore_53_14_0.xasm -- per cycle one of sqrt-X / sqrt-Y / sqrt-W on every qubit
    (never the same gate twice in a row on a qubit), then fSim(theta ~ pi/2, phi ~ pi/6) with per-coupler angles on one of
    the four coupler patterns in the order ABCDCDAB.  Qubits are numbered row-major on a rows x cols grid, so the horizontal
    couplers are nearest neighbours and the vertical ones are `cols` apart (the resource file has distances 1..10);
    nearest_neighbor() routes them exactly as TNQVM's pass would."""
    rng = np.random.Generator(np.random.PCG64(seed))
    idx = lambda r, c_: r * cols + c_
    pat = {"A": [(idx(r, c_), idx(r, c_ + 1)) for r in range(rows) for c_ in range(0, cols - 1, 2)],
           "B": [(idx(r, c_), idx(r, c_ + 1)) for r in range(rows) for c_ in range(1, cols - 1, 2)],
           "C": [(idx(r, c_), idx(r + 1, c_)) for r in range(0, rows - 1, 2) for c_ in range(cols)],
           "D": [(idx(r, c_), idx(r + 1, c_)) for r in range(1, rows - 1, 2) for c_ in range(cols)]}
    angles = {}
    for k in "ABCD":
        pat[k] = [(a, b) for (a, b) in pat[k] if a < n and b < n]
        for e in pat[k]:
            angles[e] = (math.pi / 2 + float(rng.uniform(-0.06, 0.06)), math.pi / 6 + float(rng.uniform(-0.06, 0.06)))
    last = [-1] * n
    c = []
    for cycle in range(depth):
        for q in range(n):
            k = int(rng.integers(0, 3))
            if k == last[q]:
                k = (k + 1 + int(rng.integers(0, 2))) % 3
            last[q] = k
            if k == 0:
                c.append(("Rx", (q,), (math.pi / 2,)))
            elif k == 1:
                c.append(("Ry", (q,), (math.pi / 2,)))
            else:
                c += [("Rz", (q,), (-math.pi / 4,)), ("Rx", (q,), (math.pi / 2,)), ("Rz", (q,), (math.pi / 4,))]
        for e in pat["ABCDCDAB"[cycle % 8]]:
            c.append(("fSim", e, angles[e]))
    return c


# ------------------------------------------------------------------ XASM subset
_GATE_RE = re.compile(r"^\s*([A-Za-z0-9_]+)\s*\((.*)\)\s*;\s*$")


_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.Pow: operator.pow}


def _eval_param(expr):
    """Gate parameter of an XASM line: numbers, `pi`, + - * / ** and parentheses only (a whitelisted walk over the
    parsed expression -- never eval(): a .xasm file is untrusted input)."""
    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            return float(node.value)
        if isinstance(node, ast.Name) and node.id in ("pi", "PI", "M_PI"):
            return math.pi
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.UAdd, ast.USub)):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
            return float(_BINOPS[type(node.op)](ev(node.left), ev(node.right)))
        raise ValueError("unsupported expression in XASM gate parameter: %r" % expr)
    if len(expr) > 200:
        raise ValueError("XASM gate parameter too long")
    try:
        return ev(ast.parse(expr, mode="eval"))
    except SyntaxError:
        raise ValueError("cannot parse XASM gate parameter: %r" % expr) from None


def load_xasm(text):
    """Parse the XASM subset used by the reference's tests and examples/sycamore/resources/*.xasm:
    one instruction per line, ``Gate(q[i](, q[j])(, param)*);``.  Returns (n_qubits_seen, circuit)."""
    circ = []
    nq = 0
    for line in text.splitlines():
        line = line.split("//")[0].strip()
        if not line or line.startswith("__qpu__") or line in ("{", "}"):
            continue
        m = _GATE_RE.match(line)
        if not m:
            continue
        name, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
        qs, ps = [], []
        for a in args:
            mq = re.match(r"^[A-Za-z_]\w*\[(\d+)\]$", a)
            if mq:
                qs.append(int(mq.group(1)))
            elif a:
                ps.append(_eval_param(a))
        if name == "CX":
            name = "CNOT"
        if qs:
            nq = max(nq, max(qs) + 1)
        circ.append((name, tuple(qs), tuple(ps)))
    return nq, circ


def load_circuit_fixture(path):
    """A circuit stored as data (tests/golden/circuits/*.json.gz, written by tests/golden/make_sycamore_fixture.py from the
    reference's resource files): returns (n_qubits, circuit)."""
    import gzip
    import json
    with gzip.open(path, "rb") as f:
        doc = json.loads(f.read().decode())
    return int(doc["n_qubits"]), [(g[0], tuple(int(q) for q in g[1]), tuple(float(p) for p in g[2])) for g in doc["circuit"]]


def sycamore_53(depth=14):
    """BASELINE config 5: the reference's examples/sycamore/resources/sycamore_53_<depth>_0.xasm (53 qubits, Rx/Ry/Rz + fSim,
    coupler distances 1..10), from the committed fixture; route it with nearest_neighbor()."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "circuits",
                        "sycamore_53_%d_0.json.gz" % depth)
    return load_circuit_fixture(path)


def to_xasm(circuit, name="kernel"):
    out = ["__qpu__ void %s(qbit q) {" % name]
    for g in circuit:
        nm, qs = g[0], g[1]
        ps = g[2] if len(g) > 2 else ()
        args = ["q[%d]" % q for q in qs] + [repr(float(p)) for p in ps]
        out.append("%s(%s);" % ("CX" if nm == "CNOT" else nm, ", ".join(args)))
    out.append("}")
    return "\n".join(out) + "\n"


# ------------------------------------------------------------------ nearest-neighbour rewrite
def nearest_neighbor(circuit, max_distance=1):
    """Meet-in-the-middle Swap ladders around every 2q gate with |q0-q1| > max_distance
    (NearestNeighborTransform.hpp:43-135, same swap order)."""
    out = []
    for g in circuit:
        name, qs = g[0], g[1]
        ps = g[2] if len(g) > 2 else ()
        if len(qs) == 2 and abs(qs[0] - qs[1]) > max_distance:
            lo0, hi0 = min(qs), max(qs)
            lo, hi = lo0, hi0
            while True:
                out.append(("Swap", (lo, lo + 1), ()))
                lo += 1
                if abs(lo - hi) <= max_distance:
                    break
                out.append(("Swap", (hi, hi - 1), ()))
                hi -= 1
                if abs(lo - hi) <= max_distance:
                    break
            out.append((name, (lo, hi) if qs[0] < qs[1] else (hi, lo), ps))
            for i in range(lo, lo0, -1):
                out.append(("Swap", (i, i - 1), ()))
            for i in range(hi, hi0):
                out.append(("Swap", (i, i + 1), ()))
        else:
            out.append((name, tuple(qs), tuple(ps)))
    return out


def count_gates(circuit):
    n1 = sum(1 for g in circuit if len(g[1]) == 1 and g[0] != "Measure")
    n2 = sum(1 for g in circuit if len(g[1]) == 2)
    return n1, n2
