"""Site-sharded handle (mps_create_sharded, one process, N GPUs) against the single-GPU engine on the same circuit:
parity of <Z_k>, norm, bond dimensions; wall time of both.  Prints one JSON line.
usage: python scripts/sharded_abi_check.py --gpus 2 --qubits 50 --depth 20 --chi 256 [--circuit brickwork|qaoa|sycamore]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tnqvm_b200
from tnqvm_b200 import circuits as Cc

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--qubits", type=int, default=50)
ap.add_argument("--depth", type=int, default=20)
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--circuit", default="brickwork")
ap.add_argument("--partition-by", default="cost")
ap.add_argument("--seed", type=int, default=12345)
ap.add_argument("--repeat", type=int, default=2)
a = ap.parse_args()
if a.circuit == "brickwork":
    n = a.qubits; circ = Cc.brickwork(n, a.depth, seed=a.seed)
elif a.circuit == "qaoa":
    n = a.qubits; circ = Cc.nearest_neighbor(Cc.qaoa_ring(n, a.depth, seed=7))
else:
    n, raw = Cc.sycamore_53(14); circ = Cc.nearest_neighbor(raw)
cc = tnqvm_b200.CompiledCircuit(circ)
n2 = Cc.count_gates(circ)[1]

def run(devices):
    e = tnqvm_b200.B200MPS(n, max_bond=a.chi, devices=devices, partition_by=a.partition_by)
    best = 1e30
    for r in range(a.repeat + 1):   # first pass warms allocators / lazy module loading
        e.reset()
        e.sync()
        t0 = time.perf_counter()
        e.run(cc); e.sync()
        dt = time.perf_counter() - t0
        if r:
            best = min(best, dt)
    z = e.expval_z_all(); nr = e.norm(); b = np.asarray(e.bond_dims()); st = e.stats(); lay = e.shard_layout()
    e.close()
    return best, z, nr, b, st, lay

t1, z1, n1, b1, s1, _ = run([0])
tN, zN, nN, bN, sN, lay = run(list(range(a.gpus)))
print(json.dumps({"circuit": a.circuit, "qubits": n, "chi": a.chi, "gates_2q": n2, "gpus": a.gpus, "partition_by": a.partition_by, "layout": lay,
                  "wall_ms_1gpu": t1 * 1e3, "wall_ms_sharded": tN * 1e3, "speedup": t1 / tN,
                  "max_abs_dz": float(np.abs(z1 - zN).max()), "rel_dnorm": abs(n1 - nN) / abs(n1), "bond_mismatch": int((b1 != bN).sum()),
                  "boundary_exchanges": sN["boundary_exchanges"], "peer_mbytes": sN["peer_bytes"] / 1e6,
                  "layers_1gpu": s1["layers"], "layers_sharded_sum": sN["layers"]}))
