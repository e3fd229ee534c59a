"""Developer check on a GPU box: parity vs the oracle with verbose numbers + rough timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqvm_b200
from tnqvm_b200 import circuits as Cc
from oracle import oracle as O

def check(n, circ, max_bond=0, tag="", **opt):
    t0 = time.time()
    e = tnqvm_b200.B200MPS(n, max_bond=max_bond, **opt)
    e.run(circ); e.sync()
    t1 = time.time()
    o = O.OracleMPS(n, max_bond=max_bond).run(circ)
    t2 = time.time()
    out = dict(tag=tag, n=n, gpu_s=round(t1 - t0, 3), cpu_s=round(t2 - t1, 3))
    out["bonds_gpu"] = e.bond_dims().tolist()[: 12]
    out["bonds_cpu"] = o.bond_dims().tolist()[: 12]
    out["norm"] = (e.norm(), o.norm())
    if n <= 16:
        sv, svo = e.statevector(), o.statevector()
        out["sv_err"] = float(np.abs(sv - svo).max())
    z = e.expval_z_all()
    zo = np.array([o.expval_z([k]) for k in range(n)])
    out["z_err"] = float(np.abs(z - zo).max())
    a, b2 = min(1, n - 1), n - 1
    out["zz"] = (float(e.expval_zz_pairs([(0, 1), (a, b2)])[1]), o.expval_z([a, b2]) if a != b2 else o.norm())
    out["stats"] = e.stats()
    print(out, flush=True)
    e.close()

check(2, [("H", (0,), ()), ("CNOT", (0, 1), ())], tag="bell")
check(4, Cc.ghz(4), tag="ghz4")
check(4, [("H", (1,), ()), ("CNOT", (1, 0), ()), ("X", (3,), ()), ("CNOT", (3, 2), ())], tag="rev")
check(10, Cc.brickwork(10, 8, seed=3, prefix_ghz=True), tag="bw10")
check(10, Cc.brickwork(10, 8, seed=3, prefix_ghz=True), tag="bw10 nofuse", fuse_1q=0)
check(10, Cc.brickwork(10, 8, seed=3, prefix_ghz=True), tag="bw10 seq", layer_batch=0)
check(12, Cc.brickwork(12, 10, seed=5, two_qubit="fSim"), tag="bw12 fsim(0,0)->id")
check(16, Cc.brickwork(16, 10, seed=12345, prefix_ghz=True), max_bond=64, tag="C1")
check(24, Cc.brickwork(24, 12, seed=7), max_bond=32, tag="bw24 chi32")
check(30, Cc.brickwork(30, 16, seed=9), max_bond=128, tag="bw30 chi128")
