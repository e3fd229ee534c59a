"""Generates tests/golden/*.json from the REFERENCE'S OWN dense simulator (oracle/_ref, compiled from
/root/reference/tnqvm/{base,utils} headers).  Run in the build container only:  python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tnqvm_b200 import circuits as Cc  # noqa: E402

assert O.ref() is not None, "needs oracle/_ref (reference headers)"
HERE = os.path.dirname(os.path.abspath(__file__))
cases = {
    "ghz_brickwork_n12_d8_seed12345": (12, Cc.brickwork(12, 8, seed=12345, prefix_ghz=True)),
    "brickwork_n14_d10_seed7": (14, Cc.brickwork(14, 10, seed=7)),
    "rcs_n10_l6_seed3": (10, [g for g in Cc.rcs(10, 6, seed=3) if g[0] != "Measure"]),
    "hea_n12_l3_seed1": (12, Cc.hea(12, 3, seed=1)),
}
rng = np.random.default_rng(0)
for name, (n, circ) in cases.items():
    # the reference dense simulator only has 1q gates and CNOT: all fixture circuits use exactly those
    assert all(len(g[1]) == 1 or g[0] == "CNOT" for g in circ)
    st = O.dense_run(n, circ, use_ref=True)
    amps = []
    for _ in range(8):
        idx = int(rng.integers(0, 1 << n))
        bits = [(idx >> q) & 1 for q in range(n)]
        amps.append([bits, [float(st[idx].real), float(st[idx].imag)]])
    gold = dict(n=n, circuit=[[g[0], list(g[1]), list(g[2]) if len(g) > 2 else []] for g in circ],
                expz=[O.dense_expval_z(st, n, [q]) for q in range(n)], amplitudes=amps, norm=float(np.vdot(st, st).real),
                source="reference dense simulator: tnqvm/utils/GateMatrixAlgebra.hpp ApplySingleQubitGate/ApplyCNOTGate + tnqvm/base/Gates.hpp")
    json.dump(gold, open(os.path.join(HERE, name + ".json"), "w"))
    print("wrote", name)
