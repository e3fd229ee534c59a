import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqvm_b200
from tnqvm_b200 import circuits as Cc
from oracle import oracle as O

def contract(sites):
    cur = sites[0].reshape(2, -1)
    for s in sites[1:]:
        dl, _, dr = s.shape
        cur = (cur @ s.reshape(dl, 2 * dr, order="F")).reshape(-1, dr, order="F")
    return cur.reshape(-1)

n = 16
circ = Cc.brickwork(n, 10, seed=12345, prefix_ghz=True)
e = tnqvm_b200.B200MPS(n, max_bond=64)
o = O.OracleMPS(n, max_bond=64)
cnt = 0
for g in circ:
    e.apply(*g); o.apply(*g)
    if len(g[1]) == 2:
        cnt += 1
        if cnt >= 15 and (cnt - 15) % 4 == 0:
            svo = o.statevector()
            sites = [e.get_site(k) for k in range(n)]
            sv_sites = contract(sites)
            sv_api = e.statevector()
            print(cnt, "bonds", e.bond_dims().tolist(), "sites-vs-oracle %.2e  api-vs-sites %.2e  norm api %.6f sites %.6f oracle %.6f" % (
                np.abs(sv_sites - svo).max(), np.abs(sv_api - sv_sites).max(), e.norm(), np.vdot(sv_sites, sv_sites).real, o.norm()), flush=True)
