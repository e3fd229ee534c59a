"""Developer check: one batched layer of 2q gates on random sites; per-gate reconstruction error."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqvm_b200
from tnqvm_b200.gates import gate_matrix

rng = np.random.default_rng(0)
def rnd(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(2 * s[0] * s[-1])

def layer(n, chi, start=0, max_bond=0):
    e = tnqvm_b200.B200MPS(n, max_bond=max_bond)
    dims = [1] + [chi] * (n - 1) + [1]
    S = [rnd(dims[k], 2, dims[k + 1]) for k in range(n)]
    for k in range(n): e.set_site(k, S[k])
    m = gate_matrix("fSim", (0.7, 0.3))
    pairs = [(j, j + 1) for j in range(start, n - 1, 2)]
    for (a, b) in pairs: e.apply_2q(a, b, m)
    e.sync()
    g = m.reshape(2, 2, 2, 2)
    errs = []
    for (a, b) in pairs:
        A, B = e.get_site(a), e.get_site(b)
        D = np.einsum('apk,kqc->apqc', S[a], S[b])
        th = np.einsum('pqij,aijc->apqc', g, D)
        if max_bond:
            cl, cr = th.shape[0], th.shape[3]
            U, sv_, Vh = np.linalg.svd(th.reshape(2 * cl, 2 * cr), full_matrices=False)
            kk = min(max_bond, len(sv_))
            th = ((U[:, :kk] * sv_[:kk]) @ Vh[:kk]).reshape(cl, 2, 2, cr)
        got = np.einsum('apk,kqc->apqc', A, B)
        errs.append(float(np.abs(got - th).max()))
    print("n", n, "chi", chi, "start", start, "max_bond", max_bond, "errs", ["%.1e" % x for x in errs], e.stats()["jacobi_sweeps"], flush=True)
    e.close()

layer(4, 32); layer(4, 64); layer(6, 64); layer(8, 64, 1); layer(8, 64, 0, 64); layer(6, 128, 0, 128); layer(8, 33)
