"""Gate name -> matrix for the host side (restates tnqvm/base/Gates.hpp:132-334 and the dispatch of
tnqvm/visitors/exatn-mps/ExatnUtils.cpp:80-113; row index = output).  Unknown names give the
identity like the reference (ExatnUtils.cpp:112)."""
import cmath
import math

import numpy as np

_S2 = math.sqrt(0.5)


def gate_matrix(name, params=()):
    p = list(params) + [0.0, 0.0, 0.0]
    if name == "CX":
        name = "CNOT"
    if name == "U3":
        name = "U"
    one = {
        "I": lambda: [[1, 0], [0, 1]],
        "H": lambda: [[_S2, _S2], [_S2, -_S2]],
        "X": lambda: [[0, 1], [1, 0]],
        "Y": lambda: [[0, -1j], [1j, 0]],
        "Z": lambda: [[1, 0], [0, -1]],
        "Rx": lambda: [[math.cos(0.5 * p[0]), -1j * math.sin(0.5 * p[0])], [-1j * math.sin(0.5 * p[0]), math.cos(0.5 * p[0])]],
        "Ry": lambda: [[math.cos(0.5 * p[0]), -math.sin(0.5 * p[0])], [math.sin(0.5 * p[0]), math.cos(0.5 * p[0])]],
        "Rz": lambda: [[cmath.exp(-0.5j * p[0]), 0], [0, cmath.exp(0.5j * p[0])]],
        "S": lambda: [[1, 0], [0, 1j]],
        "Sdg": lambda: [[1, 0], [0, -1j]],
        "T": lambda: [[1, 0], [0, cmath.exp(0.25j * math.pi)]],
        "Tdg": lambda: [[1, 0], [0, cmath.exp(-0.25j * math.pi)]],
        "U": lambda: [[math.cos(p[0] / 2), -cmath.exp(1j * p[2]) * math.sin(p[0] / 2)],
                      [cmath.exp(1j * p[1]) * math.sin(p[0] / 2), cmath.exp(1j * (p[1] + p[2])) * math.cos(p[0] / 2)]],
    }
    two = {
        "CNOT": lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]],
        "CZ": lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]],
        "CY": lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]],
        "CH": lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, _S2, _S2], [0, 0, _S2, -_S2]],
        "CRZ": lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, cmath.exp(-0.5j * p[0]), 0], [0, 0, 0, cmath.exp(0.5j * p[0])]],
        "CPhase": lambda: [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, cmath.exp(1j * p[0])]],
        "Swap": lambda: [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]],
        "iSwap": lambda: [[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]],
        "fSim": lambda: [[1, 0, 0, 0], [0, math.cos(p[0]), -1j * math.sin(p[0]), 0], [0, -1j * math.sin(p[0]), math.cos(p[0]), 0],
                         [0, 0, 0, cmath.exp(-1j * p[1])]],
    }
    if name in one:
        return np.array(one[name](), dtype=np.complex128)
    if name in two:
        return np.array(two[name](), dtype=np.complex128)
    return np.eye(2, dtype=np.complex128)


TWO_QUBIT = {"CNOT", "CX", "CZ", "CY", "CH", "CRZ", "CPhase", "Swap", "iSwap", "fSim"}
