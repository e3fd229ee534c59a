"""Writes tests/golden/circuits/sycamore_53_14_0.json.gz from the reference's own resource file
(/root/reference/examples/sycamore/resources/sycamore_53_14_0.xasm, the circuit BASELINE.json config 5 names).

The fixture is the instruction list of that file as data -- [name, [qubits], [params]] per gate, long-range fSim gates
untouched -- so that the GPU box (which has no /root/reference) can run the real 53-qubit depth-14 circuit:
tnqvm_b200.circuits.load_circuit_fixture() reads it, circuits.nearest_neighbor() routes it exactly as TNQVM's pass
would (NearestNeighborTransform.hpp:43-135: 301 fSim + 2 x 798 Swap = 1897 nearest-neighbour 2q gates).
Run in the build container only:  python tests/golden/make_sycamore_fixture.py
"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tnqvm_b200 import circuits as Cc   # noqa: E402

SRC = "/root/reference/examples/sycamore/resources/sycamore_53_%d_0.xasm"


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "circuits")
    os.makedirs(out_dir, exist_ok=True)
    for depth in (14,):
        n, circ = Cc.load_xasm(open(SRC % depth).read())
        assert n == 53
        n1, n2 = Cc.count_gates(circ)
        nn = Cc.nearest_neighbor(circ)
        doc = {"source": "examples/sycamore/resources/sycamore_53_%d_0.xasm" % depth, "n_qubits": n, "gates_1q": n1, "gates_2q": n2,
               "nn_gates_2q": Cc.count_gates(nn)[1],
               "circuit": [[g[0], list(g[1]), [float(p) for p in g[2]]] for g in circ]}
        path = os.path.join(out_dir, "sycamore_53_%d_0.json.gz" % depth)
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(doc, separators=(",", ":")).encode())
        print(path, os.path.getsize(path), "bytes;", n1, "1q gates,", n2, "2q gates,", doc["nn_gates_2q"], "after the nearest-neighbour pass")


if __name__ == "__main__":
    main()
