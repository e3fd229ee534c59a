import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqvm_b200
from tnqvm_b200 import circuits as Cc
from tnqvm_b200.gates import gate_matrix
from oracle import oracle as O

def contract(sites):
    cur = sites[0].reshape(2, -1)
    for s in sites[1:]:
        dl, _, dr = s.shape
        cur = (cur @ s.reshape(dl, 2 * dr, order="F")).reshape(-1, dr, order="F")
    return cur.reshape(-1)

n = 16
circ = Cc.brickwork(n, 10, seed=12345, prefix_ghz=True)
e = tnqvm_b200.B200MPS(n, max_bond=64, fuse_1q=0, layer_batch=0)
o = O.OracleMPS(n, max_bond=64)
cnt = 0
for g in circ:
    two = len(g[1]) == 2
    if two:
        cnt += 1
        lo = min(g[1])
        if 44 <= cnt <= 60:
            A0, B0 = e.get_site(lo), e.get_site(lo + 1)
    e.apply(*g); o.apply(*g)
    if two and 44 <= cnt <= 60:
        A1, B1 = e.get_site(lo), e.get_site(lo + 1)
        D = np.einsum('apk,kqc->apqc', A0, B0)
        m = gate_matrix(g[0], g[2]).reshape(2, 2, 2, 2)
        th = np.einsum('pqij,aijc->apqc', m, D) if g[1][0] == lo else np.einsum('qpji,aijc->apqc', m, D)
        got = np.einsum('apk,kqc->apqc', A1, B1)
        sv_sites = contract([e.get_site(k) for k in range(n)])
        s = np.linalg.svd(th.reshape(2 * A0.shape[0], -1), compute_uv=False)
        print(cnt, g[0], g[1], "dims", A0.shape, B0.shape, "->", A1.shape[2], "recon %.2e" % np.abs(got - th).max(),
              "state err %.2e" % np.abs(sv_sites - o.statevector()).max(), "|A0|max %.1e |B0|max %.1e" % (np.abs(A0).max(), np.abs(B0).max()),
              "sv[-3:]", s[-3:], flush=True)
